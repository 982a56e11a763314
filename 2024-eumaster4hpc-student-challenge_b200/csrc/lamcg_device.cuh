// lamcg_device.cuh — device-side state, PTX wrappers and small reductions shared by all kernels.
// sm_100a only (TMA bulk copies, mbarrier, L2 cache-policy hints).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lamcgk {

// Everything one CG solve keeps on the device so that an iteration never visits the host.
// The reference keeps alpha/beta/rr in device memory too (GPU/local/ConjugateGradient_GPU_CUDA.cu:246-254)
// but copies rr and bb back every iteration for the stop test (:285-287); here the stop test,
// the iteration counter and the `done` latch live next to them.
//
// rr[] and iter[] are double-buffered by iteration parity: inside one kernel every thread reads
// slot [par] while a single thread writes slot [par^1], so no kernel both reads and writes the
// same word and no extra "scalar" kernels (the reference's divide<<<1,1>>>, :273,:281) are needed.
struct DevState {
    double bb;         // rhs_module = b.b                      (OMP.hpp:65)
    double rr[2];      // r.r entering the iteration, by parity (OMP.hpp:67,76)
    double pAp_local;  // this rank's  sum p_i (Ap)_i  from the GEMV epilogue
    double pAp;        // all-reduced value (multi-rank only; single rank reads pAp_local)
    double rrn_local;  // this rank's  sum r_i r_i  after the update
    double rrn;        // all-reduced value
    double eps;        // rel_error
    double rr_final;   // rr after the last executed iteration
    double alpha_last, beta_last;
    int iter[2];       // iterations executed so far, by parity
    int max_iters;
    int done;          // latch: once set every later kernel of the loop is a no-op
    int converged;
    int breakdown;     // stopped on a non-finite residual / beta
    int iters_done;
    int error;         // 0 ok | 1 mbarrier timeout | 2 peer-flag timeout | 3 persistent-loop exchange timeout | 4 fused K2+K3 wait timeout
    int hist_cap;
    unsigned int ticket_gemv;
    unsigned int ticket_xr;
    unsigned int ticket_misc;
    unsigned int rrn_ready;      // fused K2+K3, single rank: iteration number whose r.r total is in rrn_local
    unsigned long long seq_base; // peer mode: flags published by this solve are seq_base + iteration index
    long long phase_cycles[8];   // persistent loop, CTA 0: SM cycles spent per phase (see lamcg_get_loop_profile)
};

// ---------------------------------------------------------------------------------------------
// Peer exchange ("comm_mode peer"): every rank owns one exchange buffer in its HBM, mapped into all
// other ranks' address spaces (CUDA IPC over NVLink/NVSwitch).  A producer kernel stores its data
// straight into every consumer's buffer, fences at system scope and then raises a sequence-numbered
// flag there; the consumer kernel spins on flags in its OWN memory.  That replaces the three
// collectives of an iteration (all-gather of p, two scalar all-reduces) by stores fused into K1/K2/K3.
// Scalars are exchanged as all-gather-of-partials + identical fixed-order local sum, so every rank
// computes bit-identical alpha/beta and trips the `done` latch on the same iteration.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;

struct PeerHeader {
    unsigned long long p_flag[kMaxRanks];    // [src] = seq: src's slice of p for iteration index (seq - seq_base) has landed
    unsigned long long pap_flag[kMaxRanks];  // [src] = seq: src's partial of p.Ap for that iteration has landed
    unsigned long long rrn_flag[kMaxRanks];
    unsigned long long gather_flag[kMaxRanks];
    double pap_slot[2][kMaxRanks];           // [iteration parity][src]
    double rrn_slot[2][kMaxRanks];
};

struct PeerView {
    unsigned char *base[kMaxRanks]; // exchange buffer of every rank as mapped in THIS process
    long long off_p[2];             // byte offsets of the two full-length p buffers
    long long off_xg[2];            // byte offsets of the two solution-gather buffers
    long long timeout_cycles;       // bound of every flag wait (peer_wait_all)
    int me, nranks;
};

__device__ __forceinline__ PeerHeader *peer_hdr(const PeerView &pv, int r) { return reinterpret_cast<PeerHeader *>(pv.base[r]); }
template <typename T = double>
__device__ __forceinline__ T *peer_p(const PeerView &pv, int r, int buf) { return reinterpret_cast<T *>(pv.base[r] + pv.off_p[buf]); }
template <typename T = double>
__device__ __forceinline__ T *peer_xg(const PeerView &pv, int r, int buf) { return reinterpret_cast<T *>(pv.base[r] + pv.off_xg[buf]); }

__device__ __forceinline__ int ld_volatile_int(const int *p) { return *(const volatile int *)p; }

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Spin until all nranks flags reach `seq`.  Bounded by pv.timeout_cycles (option peer_timeout_s, default 600 s: ranks may
// finish a cold-cache file ingest minutes apart; the NCCL path and the reference's MPI drivers would simply wait).  On timeout
// the kernel does NOT trap (a trap poisons the CUDA context of the whole process): it records error 2, latches `done` so that
// every later kernel of the loop is a no-op, and returns false in ALL threads; the caller leaves the kernel and the host
// reports LAMCG_ERR_DEVICE.  Called by every thread of a CTA; lanes 0..nranks-1 poll, the CTA barrier releases everybody.
__device__ __forceinline__ bool peer_wait_all(const unsigned long long *flags, int nranks, unsigned long long seq, DevState *st,
                                              long long timeout_cycles)
{
    int failed = 0;
    if ((int)threadIdx.x < nranks) {
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(&flags[threadIdx.x]) < seq) {
            if (clock64() - t0 > timeout_cycles || ld_volatile_int(&st->error) != 0) {
                failed = 1;
                break;
            }
        }
        if (failed) {
            st->error = 2;
            __threadfence();
            st->done = 1;
        }
    }
    return __syncthreads_or(failed) == 0;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier.init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must end in a trap (sticky error reported by the host), never in
// a hung GPU.  ~4e9 cycles is about two seconds.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err_flag)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            *err_flag = 1;
            __threadfence_system();
            __trap();
        }
    }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// 128-bit streaming load of two doubles: read-only path, no L1 allocation, L2 policy hint.
__device__ __forceinline__ double2 ldg_stream_f64x2(const double *ptr, uint64_t policy)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(ptr), "l"(policy));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Element-type helpers.  The hot path is fp64 (what the reference's drivers instantiate); the <float>
// instantiation of the reference classes is served by the same kernels with T = float for everything that
// is STORED (A, b, x, r, p, Ap) while every reduction, alpha, beta and the stop test stay in fp64.
// ---------------------------------------------------------------------------------------------
template <typename T, int VB> struct VecN { T v[VB / (int)sizeof(T)]; }; // one VB-byte load worth of elements (VB = 16 or 32)

// Streaming load of the matrix: read-only path, no L1 allocation, L2 policy hint.  VB = 32 is the 256-bit global load that
// sm_100 added (PTX .v4.f64 / .v8.f32; SASS LDG.E.NA.ENL2.256.CONSTANT).
template <typename T, int VB> __device__ __forceinline__ VecN<T, VB> ldg_stream_vec(const T *ptr, uint64_t policy);
template <> __device__ __forceinline__ VecN<double, 16> ldg_stream_vec<double, 16>(const double *ptr, uint64_t policy)
{
    VecN<double, 16> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(ptr), "l"(policy));
    return r;
}
template <> __device__ __forceinline__ VecN<double, 32> ldg_stream_vec<double, 32>(const double *ptr, uint64_t policy)
{
    VecN<double, 32> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(ptr), "l"(policy));
    return r;
}
template <> __device__ __forceinline__ VecN<float, 16> ldg_stream_vec<float, 16>(const float *ptr, uint64_t policy)
{
    VecN<float, 16> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                 : "l"(ptr), "l"(policy));
    return r;
}
template <> __device__ __forceinline__ VecN<float, 32> ldg_stream_vec<float, 32>(const float *ptr, uint64_t policy)
{
    VecN<float, 32> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(ptr), "l"(policy));
    return r;
}
// p: read-only path WITH L1 allocation (every CTA re-reads it once per pass)
template <typename T, int VB> __device__ __forceinline__ VecN<T, VB> ldg_vec(const T *ptr);
template <> __device__ __forceinline__ VecN<double, 16> ldg_vec<double, 16>(const double *ptr)
{
    const double2 t = __ldg(reinterpret_cast<const double2 *>(ptr));
    VecN<double, 16> r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    return r;
}
template <> __device__ __forceinline__ VecN<double, 32> ldg_vec<double, 32>(const double *ptr)
{
    VecN<double, 32> r;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(ptr));
    return r;
}
template <> __device__ __forceinline__ VecN<float, 16> ldg_vec<float, 16>(const float *ptr)
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(ptr));
    VecN<float, 16> r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    r.v[2] = t.z;
    r.v[3] = t.w;
    return r;
}
template <> __device__ __forceinline__ VecN<float, 32> ldg_vec<float, 32>(const float *ptr)
{
    VecN<float, 32> r;
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(ptr));
    return r;
}
// acc + a*b in fp64: unfused like the reference for doubles; for floats the product is exact in fp64
__device__ __forceinline__ double prod_acc(double a, double b, double acc) { return __dadd_rn(__dmul_rn(a, b), acc); }
__device__ __forceinline__ double prod_acc(float a, float b, double acc) { return __dadd_rn(__dmul_rn((double)a, (double)b), acc); }
// s*x + y in the STORAGE precision, unfused (axpby of the reference with the scalar rounded to T first)
__device__ __forceinline__ double scale_add(double s, double x, double y) { return __dadd_rn(__dmul_rn(s, x), y); }
__device__ __forceinline__ float scale_add(double s, float x, float y) { return __fadd_rn(__fmul_rn((float)s, x), y); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Reductions.  All sums are evaluated in a fixed order (no floating-point atomics), so a solve is
// bit-reproducible run to run for a given (n, ranks, grid).
// Arithmetic mirrors the reference CPU build (baseline x86-64, no FMA contraction): a product is
// rounded, then added (OMP.hpp:226,258).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_add(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-level sum for the vector kernels (blockDim.x a multiple of 32, <= 1024).  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *scratch /* >= 32 doubles */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < nw; ++w) s = __dadd_rn(s, scratch[w]);
    __syncthreads();
    return s;
}

// "Last CTA finishes": every CTA stores its partial, the CTA that draws the last ticket adds the
// partials in index order and publishes the total — locally, and in peer mode into slot
// [par][me] of every rank's exchange buffer followed by a release of flag[me] = seq there.
// Called by ONE full warp per CTA.  which: 0 = p.Ap (K1), 1 = r.r (K2).
__device__ __forceinline__ void grid_sum_publish(double cta_partial, double *partials, unsigned int *ticket,
                                                 double *total_out, int lane, const PeerView *pv = nullptr, int which = 0,
                                                 int par = 0, unsigned long long seq = 0, unsigned int *ready = nullptr,
                                                 unsigned int ready_value = 0)
{
    const int G = gridDim.x, bid = blockIdx.x;
    int last = 0;
    if (lane == 0) {
        __stcg(&partials[bid], cta_partial);
        __threadfence();
        last = (atomicAdd(ticket, 1u) == (unsigned)(G - 1));
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence();
        double s = 0.0;
        for (int i = lane; i < G; i += 32) s = __dadd_rn(s, __ldcg(&partials[i]));
        s = warp_sum(s);
        if (lane == 0) {
            *total_out = s;
            *ticket = 0u;
            if (ready) st_release_gpu_u32(ready, ready_value); // fused K2+K3: the CTAs of this grid are spinning on it
        }
        if (pv && pv->nranks > 1 && lane < pv->nranks) {
            PeerHeader *dst = peer_hdr(*pv, lane);
            double *slot = which == 0 ? &dst->pap_slot[par][pv->me] : &dst->rrn_slot[par][pv->me];
            unsigned long long *flag = which == 0 ? &dst->pap_flag[pv->me] : &dst->rrn_flag[pv->me];
            *reinterpret_cast<volatile double *>(slot) = s;
            __threadfence_system();
            st_release_sys_u64(flag, seq);
        }
    }
}

// Sum of the nranks partials sitting in this rank's own exchange buffer, in rank order (identical on
// every rank).  Call after peer_wait_all on the matching flags.
__device__ __forceinline__ double peer_sum_slots(const double *slots, int nranks)
{
    double s = 0.0;
    for (int r = 0; r < nranks; ++r) s = __dadd_rn(s, ld_relaxed_sys_f64(&slots[r]));
    return s;
}

} // namespace lamcgk
