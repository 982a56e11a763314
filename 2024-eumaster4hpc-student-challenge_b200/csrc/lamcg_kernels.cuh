// lamcg_kernels.cuh — the CG hot path as hand-written sm_100a kernels.
//
//   K1  gemv_tma_kernel / gemv_ldg_kernel   Ap = A_local * p  (+ fused  p.Ap  partial)   HBM-bound
//   K2  update_xr_kernel                     x += alpha p ; r -= alpha Ap ; r.r partial
//   K3  update_p_kernel                      beta, stop test, p = r + beta p
//       init_solve_kernel, generate_matrix_kernel, stream_read_kernel
//
// Reference functions replaced (all under /root/reference/challenge/main/LAM/src):
//   gemv      CPU/ConjugateGradient_CPU_OMP.hpp:246-263, GPU/local/ConjugateGradient_GPU_CUDA.cu:170-223
//   dot       OMP.hpp:219-231, GPU_CUDA.cu:64-114 (partialDot + reduce + cudaMalloc per call)
//   axpby     OMP.hpp:233-244, GPU_CUDA.cu:130-168 (axpy / minusaxpy / xpby), divide :16-20
//   stop test OMP.hpp:77, GPU_CUDA.cu:283-287 (two D2H copies + host sqrt per iteration)
//   generator CPU/ConjugateGradient_CPU_MPI_OMP.hpp:237-247 (host loop + H2D in the GPU variants)
#pragma once

#include "lamcg_device.cuh"

namespace lamcgk {

struct GemvArgs {
    const void *A;      // [rows][lda] row block of this rank (element type T of the handle), lda % 16 == 0, pad columns zero
    const void *p;      // [lda] full direction vector, zero padded
    void *Ap;           // [rows]
    double *partials;   // [grid] per-CTA partials of p.Ap
    DevState *st;
    long long rows;     // local rows
    long long lda;      // padded columns
    long long row_offset; // global index of local row 0 (p is indexed globally)
    int check_done;     // 1 inside the solve loop, 0 for the standalone GEMV hook
    int par;            // iteration parity (scalar double-buffer slot, peer slot / p buffer)
    PeerView pv;        // pv.nranks <= 1: no peer exchange
};

// Peer mode prologue of K1: p for iteration `it` is complete once every rank's K3 of iteration it-1
// has raised p_flag (the first iteration reads the locally initialised p = b).  Returns the sequence
// number this iteration publishes its scalars with.  Called by all threads of the CTA.
__device__ __forceinline__ bool gemv_peer_prologue(const GemvArgs &g, unsigned long long &seq)
{
    seq = 0ull;
    if (g.pv.nranks <= 1 || !g.check_done) return true;
    const int it = g.st->iter[g.par];
    const unsigned long long base = g.st->seq_base;
    seq = base + (unsigned long long)it + 1ull;
    // (not instrumented for option loop_profile: the extra live values cost the 128-register row sweep a spill; the wait was
    // measured once at 0.3-0.9 us per iteration on 8 GPUs, profiles/r02_mgpu_profile_n100k_8gpu.log)
    if (it >= 1) return peer_wait_all(peer_hdr(g.pv, g.pv.me)->p_flag, g.pv.nranks, base + (unsigned long long)it, g.st, g.pv.timeout_cycles);
    return true;
}

// =============================================================================================
// K1, variant 2 ("tma ring"): the whole A stream goes through TMA bulk copies.
//
// One CTA per SM, persistent over a balanced contiguous range of rows.  A producer warp keeps a
// STAGES-deep shared-memory ring full: one stage = RB row segments of CB columns (RB bulk copies
// of CB*8 bytes, L2 evict-first: A is read once per iteration and is far larger than L2) plus the
// matching CB-column slice of p (one bulk copy, L2 evict-last: p is re-read by every CTA).
// STAGES*RB*CB*8 bytes are in flight per SM regardless of register pressure or occupancy, which
// is what covers HBM latency at ~50 GB/s per SM.
// CB/32 consumer warps: thread t owns column t of the stage and all RB rows (RB accumulators), so
// p is read from shared memory once per RB elements of A and every shared-memory access is a
// conflict-free 8-byte-per-lane row.  At the end of a pass (RB rows x all columns) the RB
// accumulators are reduced by warp shuffles, then across warps in fixed order, written to Ap, and
// p[row]*Ap[row] is added to the CTA's partial of the fused dot product.
// =============================================================================================
template <int RB, int CB, int STAGES>
struct GemvTmaCfg {
    static constexpr int kConsumerWarps = CB / 32;
    static constexpr int kThreads = (kConsumerWarps + 1) * 32;
    static constexpr int kStageDoubles = (RB + 1) * CB;
    static constexpr size_t kSmemBytes =
        (size_t)STAGES * kStageDoubles * 8 + 2 * STAGES * 8 + (size_t)kConsumerWarps * RB * 8 + 64;
};

template <int RB, int CB, int STAGES>
__global__ void __launch_bounds__(GemvTmaCfg<RB, CB, STAGES>::kThreads, 1) lamcg_tmaring_kernel(GemvArgs g)
{
    const double *gA = static_cast<const double *>(g.A), *gp = static_cast<const double *>(g.p);
    double *gAp = static_cast<double *>(g.Ap);
    using Cfg = GemvTmaCfg<RB, CB, STAGES>;
    constexpr int NW = Cfg::kConsumerWarps;
    static_assert(RB <= 32, "final reduction assumes RB <= 32");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * Cfg::kStageDoubles * 8);
    uint64_t *empty = full + STAGES;
    double *red = reinterpret_cast<double *>(empty + STAGES); // [NW][RB]

    if (g.check_done && ld_volatile_int(&g.st->done)) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, bid = blockIdx.x;
    const long long base = g.rows / G, rem = g.rows % G;
    const long long r0 = bid * base + (bid < rem ? bid : rem);
    const long long rcnt = base + (bid < rem ? 1 : 0);
    const int npass = (int)((rcnt + RB - 1) / RB);
    const int nchunk = (int)((g.lda + CB - 1) / CB);
    const long long total = (long long)npass * nchunk;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    unsigned long long seq;
    if (!gemv_peer_prologue(g, seq)) return; // peer flag timeout: error recorded, loop latched done

    if (warp == NW) {
        // ------------------------------------------------------------------ producer warp
        // A 1-D bulk copy costs the issuing thread ~70-100 cycles (measured, profiles/r01_sweep_n100k_first.log:
        // single-thread issue capped 1-2 KB copies at 4-5 TB/s), so the RB row copies of a stage are issued
        // by RB different lanes of this warp in one go; lane 0 arms the barrier first, lane RB % 32 .. fetches p.
        const uint64_t polA = l2_policy_evict_first();
        const uint64_t polP = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        int pass = 0, k = 0;
        for (long long j = 0; j < total; ++j) {
            if (j >= STAGES) mbar_wait(&empty[s], ph ^ 1u, &g.st->error); // all lanes wait: all must see the slot free
            const long long prow = r0 + (long long)pass * RB;
            const long long left = rcnt - (long long)pass * RB;
            const int nr = left < RB ? (int)left : RB;
            const long long c0 = (long long)k * CB;
            const long long cl = g.lda - c0;
            const uint32_t seg = (uint32_t)((cl < CB ? cl : CB) * 8);
            double *tile = tiles + (size_t)s * Cfg::kStageDoubles;
            if (lane == 0) mbar_arrive_expect_tx(&full[s], seg * (uint32_t)(nr + 1));
            __syncwarp();
            const double *src = gA + prow * g.lda + c0;
            for (int r = lane; r < nr; r += 32) tma_load_1d(tile + r * CB, src + (long long)r * g.lda, seg, &full[s], polA);
            if (lane == (RB & 31)) tma_load_1d(tile + RB * CB, gp + c0, seg, &full[s], polP);
            if (++k == nchunk) { k = 0; ++pass; }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int col = warp * 32 + lane; // column inside the stage
    double cta_dot = 0.0;             // meaningful in warp 0 lane 0
    int s = 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < npass; ++pass) {
        const long long prow = r0 + (long long)pass * RB;
        const long long left = rcnt - (long long)pass * RB;
        const int nr = left < RB ? (int)left : RB;
        double acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.0;

        for (int k = 0; k < nchunk; ++k) {
            mbar_wait(&full[s], ph, &g.st->error);
            const double *tile = tiles + (size_t)s * Cfg::kStageDoubles;
            const long long cl = g.lda - (long long)k * CB;
            if (col < cl) {
                const double pv = tile[RB * CB + col];
                if (nr == RB) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) acc[r] = mul_add(tile[r * CB + col], pv, acc[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < RB; ++r)
                        if (r < nr) acc[r] = mul_add(tile[r * CB + col], pv, acc[r]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }

        // pass epilogue: RB row sums -> Ap, and the fused p.Ap contribution
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RB; ++r) red[warp * RB + r] = acc[r];
        }
        named_bar_sync(1, NW * 32);
        if (warp == 0) {
            double contrib = 0.0;
            if (lane < nr) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) sum = __dadd_rn(sum, red[w * RB + lane]);
                gAp[prow + lane] = sum;
                contrib = __dmul_rn(gp[g.row_offset + prow + lane], sum);
            }
            contrib = warp_sum(contrib);
            cta_dot = __dadd_rn(cta_dot, contrib);
        }
        named_bar_sync(1, NW * 32);
    }

    if (warp == 0) grid_sum_publish(cta_dot, g.partials, &g.st->ticket_gemv, &g.st->pAp_local, lane, &g.pv, 0, g.par, seq);
}

// =============================================================================================
// K1, variant 1 ("ldg"): A through 128-bit read-only vector loads, p staged in shared memory.
//
// 8 warps per CTA, each warp owns R consecutive rows of the current pass and sweeps the columns:
// one warp-wide load instruction covers 512 contiguous bytes of one row (4 full 128-byte lines,
// sectors/request = 4), R*U such loads are issued back to back before the first use.  p is staged
// through a 3-stage shared-memory ring of 2048-column chunks filled by TMA bulk copies (thread 0 is
// the producer, mbarrier full/empty handshakes), so p costs n*8 bytes of L2 traffic per
// (8*R)-row pass instead of per row as in the reference kernel (GPU_CUDA.cu:184-191).
// =============================================================================================
constexpr int kLdgWarps = 8;
constexpr int kLdgPChunk = 2048;
constexpr int kLdgPStages = 3;
constexpr size_t kLdgSmemBytes = (size_t)kLdgPStages * kLdgPChunk * 8 + 2 * kLdgPStages * 8 + kLdgWarps * 8 + 64;

template <int R, int U, int CPS = 2>
__global__ void __launch_bounds__(kLdgWarps * 32, CPS) lamcg_warprows_tmap_kernel(GemvArgs g)
{
    const double *gA = static_cast<const double *>(g.A), *gp = static_cast<const double *>(g.p);
    double *gAp = static_cast<double *>(g.Ap);
    constexpr int NW = kLdgWarps, PCH = kLdgPChunk, PST = kLdgPStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *pbuf = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)PST * PCH * 8);
    uint64_t *empty = full + PST;
    double *wdot = reinterpret_cast<double *>(empty + PST);

    if (g.check_done && ld_volatile_int(&g.st->done)) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, bid = blockIdx.x;
    const long long base = g.rows / G, rem = g.rows % G;
    const long long r0 = bid * base + (bid < rem ? bid : rem);
    const long long rcnt = base + (bid < rem ? 1 : 0);
    constexpr int RP = NW * R; // rows per pass
    const int npass = (int)((rcnt + RP - 1) / RP);
    const int nchunk = (int)((g.lda + PCH - 1) / PCH);
    const long long total = (long long)npass * nchunk;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < PST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    unsigned long long seq;
    if (!gemv_peer_prologue(g, seq)) return; // peer flag timeout: error recorded, loop latched done

    const uint64_t polA = l2_policy_evict_first();
    const uint64_t polP = l2_policy_evict_last();
    const bool is_producer = (threadIdx.x == 0);

    // producer side: chunk index jj -> stage jj % PST
    auto issue = [&](long long jj) {
        const int ss = (int)(jj % PST);
        const uint32_t pph = (uint32_t)((jj / PST) & 1);
        if (jj >= PST) mbar_wait(&empty[ss], pph ^ 1u, &g.st->error);
        const long long c0 = (jj % nchunk) * (long long)PCH;
        const long long cl = g.lda - c0;
        const uint32_t bytes = (uint32_t)((cl < PCH ? cl : PCH) * 8);
        mbar_arrive_expect_tx(&full[ss], bytes);
        tma_load_1d(pbuf + (size_t)ss * PCH, gp + c0, bytes, &full[ss], polP);
    };
    if (is_producer) {
        for (long long jj = 0; jj < PST - 1 && jj < total; ++jj) issue(jj);
    }

    double warp_dot = 0.0;
    long long j = 0;
    for (int pass = 0; pass < npass; ++pass) {
        const long long wrow = r0 + (long long)pass * RP + warp * R; // first row of this warp
        long long leftw = rcnt - ((long long)pass * RP + warp * R);
        const int nr = leftw <= 0 ? 0 : (leftw < R ? (int)leftw : R);
        const double *arow[R];
#pragma unroll
        for (int r = 0; r < R; ++r) arow[r] = gA + (wrow + (r < nr ? r : 0)) * g.lda;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;

        for (int k = 0; k < nchunk; ++k, ++j) {
            if (is_producer && j + PST - 1 < total) issue(j + PST - 1);
            const int s = (int)(j % PST);
            const uint32_t ph = (uint32_t)((j / PST) & 1);
            mbar_wait(&full[s], ph, &g.st->error);
            const double *pb = pbuf + (size_t)s * PCH;
            const long long c0 = (long long)k * PCH;
            const long long cl = g.lda - c0;
            const int nc = cl < PCH ? (int)cl : PCH; // multiple of 16
            if (nr > 0) {
                for (int c = 2 * lane; c < nc; c += 64 * U) {
                    double2 a[R][U], pv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int cc = c + 64 * u;
                        const bool cv = cc < nc;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (cv && r < nr) a[r][u] = ldg_stream_f64x2(arow[r] + c0 + cc, polA);
                            else a[r][u] = make_double2(0.0, 0.0);
                        }
                        pv[u] = cv ? *reinterpret_cast<const double2 *>(pb + cc) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            acc[r] = mul_add(a[r][u].x, pv[u].x, acc[r]);
                            acc[r] = mul_add(a[r][u].y, pv[u].y, acc[r]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nr) {
                    gAp[wrow + r] = acc[r];
                    warp_dot = mul_add(gp[g.row_offset + wrow + r], acc[r], warp_dot);
                }
            }
        }
    }

    if (lane == 0) wdot[warp] = warp_dot;
    __syncthreads();
    if (warp == 0) {
        double cta_dot = 0.0;
        if (lane == 0) {
#pragma unroll
            for (int w = 0; w < NW; ++w) cta_dot = __dadd_rn(cta_dot, wdot[w]);
        }
        grid_sum_publish(cta_dot, g.partials, &g.st->ticket_gemv, &g.st->pAp_local, lane, &g.pv, 0, g.par, seq);
    }
}

// =============================================================================================
// K1, default ("cta rows"): the whole CTA sweeps one row at a time, R rows in flight.
// Thread t owns E = VB/sizeof(T) consecutive columns out of every NT*E of a CH = NT*U*E column chunk: one CTA-wide load
// instruction covers NT*VB bytes of contiguous row (8 KB with 512 threads and 128-bit loads, 16 KB with the sm_100 256-bit
// form), a chunk of one row is U of those.  The thread's U vectors of p for the chunk sit in registers and are reused for the R
// rows (p is read from L2 once per R rows); R accumulators per thread are reduced across the CTA at the end of the pass.
// No shared-memory staging, no mbarriers.
// Rows are dealt to the CTAs as balanced contiguous ranges (they differ by at most one row).  A CTA's range is swept in full
// R-row passes; the remaining rows (< R) go through passes of 4, 2 and 1 rows with proportionally MORE loads per row, so every
// pass keeps the same R*U vector loads in flight per thread: at 12500 rows per GPU (n = 100000 on 8 GPUs, 84-85 rows per CTA)
// a plain "8 rows, 4 of them masked" last pass ran with half the bytes in flight for ~1/11 of the kernel.
// Mixed storage (S = float under T = double, option matrix_f32): only the MATRIX is held in fp32; every element is widened to
// fp64 (exact) before the same unfused multiply and add, so the arithmetic is the fp64 sweep's on the matrix fl32(A) and the
// HBM stream is half as long.  The widening (F2F.F64.F32) is not a limiter: 7.3 TB/s of fp32 matrix, tools/mixed_probe.cu,
// profiles/r02_mixed_probe.log.
// =============================================================================================
constexpr int kCtaRowMaxR = 8;
struct kTrue { static constexpr bool value = true; };
struct kFalse { static constexpr bool value = false; };

// One pass: R rows starting at arow0 (row stride lda), all lda columns.  Returns (warp 0, all lanes) the sum over the R rows of
// p[row] * Ap[row]; 0.0 in the other warps.  red: [NT/32][kCtaRowMaxR] shared scratch.
template <typename T, int R, int U, int NT, int VB, typename S = T>
__device__ __forceinline__ double ctarow_pass(const S *__restrict__ arow0, const T *__restrict__ gp, T *__restrict__ Ap_rs,
                                              const T *__restrict__ p_rs, long long lda, uint64_t polA,
                                              double (*red)[kCtaRowMaxR])
{
    constexpr int E = VB / (int)sizeof(S); // matrix elements per load
    constexpr int PB = E * (int)sizeof(T); // bytes of p that go with one matrix load
    constexpr int CH = NT * U * E;         // columns per chunk
    static_assert(PB == 16 || PB == 32, "p is fetched with one 128- or 256-bit load per matrix load");
    constexpr int NWARP = NT / 32;
    static_assert(R <= kCtaRowMaxR, "row sums are finished by one warp from red[][kCtaRowMaxR]");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    // One chunk: U vectors of p, then R*U streaming loads of A issued back to back, then the products.  Row r is addressed
    // as (row r-1) + lda by pointer increments and the U loads of a row by immediate offsets, so the full-chunk loop below
    // is loads + 2 integer adds per row (the first version recomputed (row)*lda in 64 bits per load under a predicate:
    // 527 instructions per chunk for 36 loads, profiles/r02_sass_k1_summary.txt).
    auto chunk = [&](const S *__restrict__ pa, const T *__restrict__ pp, auto pred, long long c_first) {
        constexpr bool kPred = decltype(pred)::value;
        VecN<T, PB> pv[U];
        bool cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cv[u] = !kPred || c_first + (long long)u * NT * E < lda; // lda % 16 == 0 and E <= 8: a vector is all in or all out
            if (cv[u]) pv[u] = ldg_vec<T, PB>(pp + u * NT * E);
            else {
#pragma unroll
                for (int e = 0; e < E; ++e) pv[u].v[e] = T(0);
            }
        }
        VecN<S, VB> a[R][U];
        const S *pr = pa;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cv[u]) a[r][u] = ldg_stream_vec<S, VB>(pr + u * NT * E, polA);
                else {
#pragma unroll
                    for (int e = 0; e < E; ++e) a[r][u].v[e] = S(0);
                }
            }
            pr += lda;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if constexpr (sizeof(T) == 4) {
                // fp32 storage: the U*E products of this thread's slice of the chunk are summed in fp32 (unfused, like the
                // reference's float loop, but only U*E terms deep), then folded into the fp64 accumulator once per chunk:
                // converting every product to fp64 made the kernel FP64-pipe bound (same 10.9 ms as the fp64 sweep for half
                // the bytes); this keeps it on the HBM roofline.
                float part = 0.0f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int e = 0; e < E; ++e) part = __fadd_rn(__fmul_rn(a[r][u].v[e], pv[u].v[e]), part);
                }
                acc[r] = __dadd_rn(acc[r], (double)part);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[r] = prod_acc((T)a[r][u].v[e], pv[u].v[e], acc[r]); // (T): exact widening when S = float
                }
            }
        }
    };
    const S *pa = arow0 + E * tid;
    const T *pp = gp + E * tid;
    const long long nfull = lda / CH;
    for (long long k = 0; k < nfull; ++k, pa += CH, pp += CH) chunk(pa, pp, kFalse{}, 0);
    if (nfull * CH < lda) chunk(pa, pp, kTrue{}, nfull * CH + E * tid);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) red[warp][r] = acc[r];
    }
    __syncthreads();
    double contrib = 0.0;
    if (warp == 0) {
        if (lane < R) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) sum = __dadd_rn(sum, red[w][lane]);
            const T stored = (T)sum;
            Ap_rs[lane] = stored;
            contrib = __dmul_rn((double)p_rs[lane], (double)stored);
        }
        contrib = warp_sum(contrib);
    }
    __syncthreads();
    return contrib;
}

template <typename T, int R, int U = 4, int NT = 256, int CPS = 2, int VB = 16, typename S = T>
__global__ void __launch_bounds__(NT, CPS) lamcg_rowsweep_kernel(GemvArgs g)
{
    static_assert(R == 8, "the tail decomposition below is written for 8-row passes");
    // widest tail pass: 4*U loads per row; with a narrower matrix type the p registers per load double, so stop at 2*U
    constexpr int UT = sizeof(S) < sizeof(T) ? 2 * U : 4 * U;
    __shared__ double red[NT / 32][kCtaRowMaxR];
    const S *gA = static_cast<const S *>(g.A);
    const T *gp = static_cast<const T *>(g.p);
    T *gAp = static_cast<T *>(g.Ap);
    if (g.check_done && ld_volatile_int(&g.st->done)) return;
    unsigned long long seq;
    if (!gemv_peer_prologue(g, seq)) return; // peer flag timeout: error recorded, loop latched done
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const long long base = g.rows / G, rem = g.rows % G;
    const long long r0 = bid * base + (bid < rem ? bid : rem);
    const long long rcnt = base + (bid < rem ? 1 : 0);
    const uint64_t polA = l2_policy_evict_first();
    const T *prow = gp + g.row_offset; // p is indexed globally
    double cta_dot = 0.0;
    long long rs = r0;
    const long long rend = r0 + rcnt;
    for (; rs + R <= rend; rs += R)
        cta_dot = __dadd_rn(cta_dot, ctarow_pass<T, R, U, NT, VB, S>(gA + rs * g.lda, gp, gAp + rs, prow + rs, g.lda, polA, red));
    // tail: < 8 rows left, same number of loads in flight per pass
    if (rs + 4 <= rend) {
        cta_dot = __dadd_rn(cta_dot, ctarow_pass<T, 4, 2 * U, NT, VB, S>(gA + rs * g.lda, gp, gAp + rs, prow + rs, g.lda, polA, red));
        rs += 4;
    }
    if (rs + 2 <= rend) {
        cta_dot = __dadd_rn(cta_dot, ctarow_pass<T, 2, UT, NT, VB, S>(gA + rs * g.lda, gp, gAp + rs, prow + rs, g.lda, polA, red));
        rs += 2;
    }
    if (rs < rend) cta_dot = __dadd_rn(cta_dot, ctarow_pass<T, 1, UT, NT, VB, S>(gA + rs * g.lda, gp, gAp + rs, prow + rs, g.lda, polA, red));
    if (warp == 0) grid_sum_publish(cta_dot, g.partials, &g.st->ticket_gemv, &g.st->pAp_local, lane, &g.pv, 0, g.par, seq);
}

// =============================================================================================
// K2: x += alpha p ; r -= alpha Ap ; partial r.r      (OMP.hpp:71-74; GPU_CUDA.cu:271-279)
// alpha = rr / (p.Ap) is recomputed by every thread from device-resident scalars.
// =============================================================================================
struct VecArgs {
    DevState *st;
    const double *pAp_src;  // &st->pAp_local (single rank) or &st->pAp (after the NCCL all-reduce)
    const double *rrn_src;  // &st->rrn_local or &st->rrn
    void *x, *r, *Ap;       // local slices [rows], element type T of the handle
    const void *p_in;       // [lda] full p of this iteration
    void *p_out;            // [lda] full p of the next iteration (== p_in except in peer mode)
    double *partials;       // [grid]
    double *hist;           // [hist_cap] sqrt(rr/bb) per iteration (nullable)
    long long rows, row_offset;
    int par;                // iteration parity for the double-buffered scalars
    int fused;              // 1: launched as update_fused_kernel (K2's last CTA also releases st->rrn_ready)
    int prof;               // 1: CTA 0 adds the SM cycles it spends per phase to st->phase_cycles[2..5] (option loop_profile)
    PeerView pv;            // pv.nranks <= 1: no peer exchange
};

// Body of K2.  Returns false (in all threads) when a peer flag wait timed out: the caller leaves the kernel.
template <typename T>
__device__ __forceinline__ bool xr_phase(const VecArgs &v, double *scratch /* 32 */, double *s_pAp)
{
    DevState *st = v.st;
    T *vx = static_cast<T *>(v.x), *vr = static_cast<T *>(v.r);
    const T *vAp = static_cast<const T *>(v.Ap);
    unsigned long long seq = 0ull;
    double pAp;
    const bool prof = v.prof && blockIdx.x == 0 && threadIdx.x == 0;
    long long tc = prof ? clock64() : 0;
    if (v.pv.nranks > 1) { // peer mode: the all-reduce of p.Ap is a wait on nranks flags + a fixed-order sum
        seq = st->seq_base + (unsigned long long)st->iter[v.par] + 1ull;
        PeerHeader *me = peer_hdr(v.pv, v.pv.me);
        if (!peer_wait_all(me->pap_flag, v.pv.nranks, seq, st, v.pv.timeout_cycles)) return false;
        if (prof) { const long long now = clock64(); st->phase_cycles[2] += now - tc; tc = now; } // wait for every rank's p.Ap
        if (threadIdx.x == 0) *s_pAp = peer_sum_slots(me->pap_slot[v.par], v.pv.nranks);
        __syncthreads();
        pAp = *s_pAp;
    } else {
        pAp = *v.pAp_src;
    }
    const double alpha = st->rr[v.par] / pAp;
    const double nalpha = -alpha;
    const T *p = static_cast<const T *>(v.p_in) + v.row_offset;
    double local = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.rows; i += (long long)gridDim.x * blockDim.x) {
        // axpby(alpha, p, 1.0, x) and axpby(-alpha, Ap, 1.0, r): alpha*x[i] + beta*y[i], unfused, in T
        vx[i] = scale_add(alpha, p[i], vx[i]);
        const T rn = scale_add(nalpha, vAp[i], vr[i]);
        vr[i] = rn;
        local = prod_acc(rn, rn, local);
    }
    const double cta = block_sum(local, scratch);
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->alpha_last = alpha;
        grid_sum_publish(cta, v.partials, &st->ticket_xr, &st->rrn_local, threadIdx.x, &v.pv, 1, v.par, seq, v.fused ? &st->rrn_ready : nullptr,
                         (unsigned int)(st->iter[v.par] + 1));
    }
    if (prof) st->phase_cycles[3] += clock64() - tc; // x, r update, r.r partial, publish
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256) update_xr_kernel(VecArgs v)
{
    __shared__ double scratch[32];
    __shared__ double s_pAp;
    if (ld_volatile_int(&v.st->done)) return;
    xr_phase<T>(v, scratch, &s_pAp);
}

// p = r + beta p on this rank's rows: the second half of K3, also run on its own by resume_kernel.
// Peer mode: the new slice goes straight into buffer [par^1] of every rank and the last CTA raises
// p_flag[me] = seq_base + it on every rank once all stores are fenced (this IS the all-gather).
template <typename T>
__device__ __forceinline__ void p_update(const VecArgs &v, double beta, int it, int *s_last)
{
    DevState *st = v.st;
    const T *vr = static_cast<const T *>(v.r);
    const T *p = static_cast<const T *>(v.p_in) + v.row_offset;
    if (v.pv.nranks > 1) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.rows; i += (long long)gridDim.x * blockDim.x) {
            const T pn = scale_add(beta, p[i], vr[i]);
            for (int q = 0; q < v.pv.nranks; ++q) peer_p<T>(v.pv, q, v.par ^ 1)[v.row_offset + i] = pn;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            *s_last = (atomicAdd(&st->ticket_misc, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (*s_last) {
            if (threadIdx.x == 0) st->ticket_misc = 0u;
            if ((int)threadIdx.x < v.pv.nranks) {
                __threadfence_system();
                st_release_sys_u64(&peer_hdr(v.pv, threadIdx.x)->p_flag[v.pv.me], st->seq_base + (unsigned long long)it);
            }
        }
    } else {
        T *po = static_cast<T *>(v.p_out) + v.row_offset;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.rows; i += (long long)gridDim.x * blockDim.x)
            po[i] = scale_add(beta, p[i], vr[i]); // axpby(1.0, r, beta, p)
    }
}

// =============================================================================================
// K3: beta = rr_new / rr ; rr = rr_new ; stop test ; p = r + beta p   (OMP.hpp:75-78)
// Every thread evaluates the (identical) scalars; thread 0 of CTA 0 commits them to the other
// parity slot and latches `done`.  In peer mode the new p slice is stored straight into every
// rank's next-iteration p buffer over NVLink (this IS the all-gather), and the last CTA raises
// p_flag on every rank once all stores are fenced.
// The p update is skipped on the final iteration (converged — as in the reference, which breaks
// before it — or max_iters reached, where nobody reads p again).
// =============================================================================================
// Body of K3.  rr_ready: the caller has already waited for r.r (fused kernel, single rank) — read it with a volatile load.
template <typename T>
__device__ __forceinline__ void p_phase(const VecArgs &v, double *s_rrn, int *s_last)
{
    DevState *st = v.st;
    const int it0 = st->iter[v.par];
    double rr_new;
    const bool prof = v.prof && blockIdx.x == 0 && threadIdx.x == 0;
    long long tc = prof ? clock64() : 0;
    if (v.pv.nranks > 1) {
        const unsigned long long seq = st->seq_base + (unsigned long long)it0 + 1ull;
        PeerHeader *me = peer_hdr(v.pv, v.pv.me);
        if (!peer_wait_all(me->rrn_flag, v.pv.nranks, seq, st, v.pv.timeout_cycles)) return;
        if (prof) { const long long now = clock64(); st->phase_cycles[4] += now - tc; tc = now; } // wait for every rank's r.r
        if (threadIdx.x == 0) *s_rrn = peer_sum_slots(me->rrn_slot[v.par], v.pv.nranks);
        __syncthreads();
        rr_new = *s_rrn;
    } else {
        rr_new = *reinterpret_cast<const volatile double *>(v.rrn_src);
    }
    const double rr_old = st->rr[v.par];
    const double beta = rr_new / rr_old;
    const int it = it0 + 1;
    const double rel = sqrt(rr_new / st->bb);
    const bool conv = rel < st->eps;
    // Non-finite scalars (b = 0, a matrix that is not SPD, overflow) can never satisfy the stop test: the
    // reference keeps iterating on NaNs until max_iters and prints "max_iters+1, nan" (TESTS/BEST_RESULTS:114).
    // Same report here, without burning the remaining iterations.
    const bool broke = !(rel == rel) || isinf(rel) || !(beta == beta);
    const bool fin = conv || broke || it >= st->max_iters;

    if (!fin) p_update<T>(v, beta, it, s_last);
    if (prof) st->phase_cycles[5] += clock64() - tc; // beta, stop test, p = r + beta p, peer stores of the slice, fence, flags
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->rr[v.par ^ 1] = rr_new;
        st->iter[v.par ^ 1] = it;
        st->iters_done = it;
        st->rr_final = rr_new;
        st->beta_last = beta;
        if (v.hist && it - 1 < st->hist_cap) v.hist[it - 1] = rel;
        if (fin) {
            st->converged = conv ? 1 : 0;
            st->breakdown = broke ? 1 : 0;
            __threadfence();
            st->done = 1;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) update_p_kernel(VecArgs v)
{
    __shared__ double s_rrn;
    __shared__ int s_last;
    if (ld_volatile_int(&v.st->done)) return;
    p_phase<T>(v, &s_rrn, &s_last);
}

// =============================================================================================
// K2 + K3 in ONE cooperative launch (single rank and peer mode; NCCL mode needs the stream between them for its all-reduce):
// x, r update and the r.r partials, then every CTA waits until r.r is complete — single rank: the CTA that drew the last ticket
// releases st->rrn_ready = iteration; peer mode: the nranks flags of the fused exchange, as K3 did at its start — and goes on with
// beta, the stop test and p = r + beta p.  One launch (and its gap on the stream) fewer per iteration; the launch is cooperative
// because every CTA spins on a value that the last CTA of the same grid produces, so all of them must be resident.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256, 4) update_fused_kernel(VecArgs v) // 4 CTAs per SM must fit: vec_grid() relies on it
{
    __shared__ double scratch[32];
    __shared__ double s_scalar;
    __shared__ int s_last;
    DevState *st = v.st;
    if (ld_volatile_int(&st->done)) return;
    if (!xr_phase<T>(v, scratch, &s_scalar)) return;
    if (v.pv.nranks <= 1) {
        const unsigned int want = (unsigned int)(st->iter[v.par] + 1);
        int failed = 0;
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(&st->rrn_ready) != want) {
                if (clock64() - t0 > 4000000000LL) { // ~2 s: the grid is co-resident (cooperative launch), so this is a bug trap, not a wait
                    failed = 1;
                    break;
                }
            }
            if (failed) {
                st->error = 4;
                __threadfence();
                st->done = 1;
            }
        }
        if (__syncthreads_or(failed)) return;
    }
    p_phase<T>(v, &s_scalar, &s_last);
}

// Continue a solve that stopped on max_iters (lamcg_solve_resume): K3 skipped the p update of its last
// iteration, so do it now with the stored beta (v is built for the parity of that last iteration), then
// raise the iteration cap and drop the `done` latch.  Iteration numbering, the scalar parity slots and the
// peer sequence numbers simply carry on, which makes solve(k) + resume(m) bit-identical to solve(k + m).
template <typename T>
__global__ void __launch_bounds__(256) resume_kernel(VecArgs v, int new_max_iters, double eps, int hist_cap)
{
    __shared__ int s_last;
    DevState *st = v.st;
    p_update<T>(v, st->beta_last, st->iters_done, &s_last);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->max_iters = new_max_iters;
        st->eps = eps;
        st->hist_cap = hist_cap;
        st->done = 0;
    }
}

// Solution gather in peer mode (lamcg_get_solution): every rank stores its x slice into buffer
// `buf` of every rank's exchange area, then raises gather_flag; the wait kernel blocks the stream
// until all slices have landed locally.
template <typename T>
__global__ void __launch_bounds__(256) peer_gather_put_kernel(PeerView pv, const void *x_, long long rows, long long row_offset,
                                                               int buf, unsigned long long gseq, DevState *st)
{
    __shared__ int s_last;
    const T *x = static_cast<const T *>(x_);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const T xi = x[i];
        for (int q = 0; q < pv.nranks; ++q) peer_xg<T>(pv, q, buf)[row_offset + i] = xi;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&st->ticket_misc, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x == 0) st->ticket_misc = 0u;
        if ((int)threadIdx.x < pv.nranks) {
            __threadfence_system();
            st_release_sys_u64(&peer_hdr(pv, threadIdx.x)->gather_flag[pv.me], gseq);
        }
    }
}

__global__ void __launch_bounds__(32) peer_gather_wait_kernel(PeerView pv, unsigned long long gseq, DevState *st)
{
    peer_wait_all(peer_hdr(pv, pv.me)->gather_flag, pv.nranks, gseq, st, pv.timeout_cycles);
}

// =============================================================================================
// Solve initialisation (OMP.hpp:56-67): x = 0, r = b_local, p = b, Ap = 0, bb = rr = b.b.
// One CTA; b.b is summed over the FULL rhs on every rank in the same fixed order, so all ranks
// hold the same bits without a collective.
// =============================================================================================
struct InitArgs {
    DevState *st;
    const void *b_full; // [lda] zero padded, element type T
    void *x, *r, *Ap, *p_full;
    long long n, lda, rows, row_offset;
    double eps;
    unsigned long long seq_base;
    int max_iters, hist_cap;
};

template <typename T>
__global__ void __launch_bounds__(1024) init_solve_kernel(InitArgs a)
{
    __shared__ double scratch[32];
    const T *b = static_cast<const T *>(a.b_full);
    T *px = static_cast<T *>(a.x), *pr = static_cast<T *>(a.r), *pAp = static_cast<T *>(a.Ap), *pp = static_cast<T *>(a.p_full);
    double local = 0.0;
    for (long long i = threadIdx.x; i < a.lda; i += blockDim.x) {
        const T bi = b[i];
        pp[i] = bi;
        local = prod_acc(bi, bi, local);
    }
    for (long long i = threadIdx.x; i < a.rows; i += blockDim.x) {
        px[i] = T(0);
        pAp[i] = T(0);
        pr[i] = b[a.row_offset + i];
    }
    const double bb = block_sum(local, scratch);
    if (threadIdx.x == 0) {
        DevState *st = a.st;
        st->bb = bb;
        st->rr[0] = bb;
        st->rr[1] = bb;
        st->iter[0] = 0;
        st->iter[1] = 0;
        st->pAp_local = st->pAp = st->rrn_local = st->rrn = 0.0;
        st->eps = a.eps;
        st->rr_final = bb;
        st->alpha_last = st->beta_last = 0.0;
        st->max_iters = a.max_iters;
        st->done = a.max_iters <= 0 ? 1 : 0;
        st->converged = 0;
        st->breakdown = 0;
        st->iters_done = 0;
        st->error = 0;
        st->hist_cap = a.hist_cap;
        st->ticket_gemv = st->ticket_xr = st->ticket_misc = 0u;
        st->rrn_ready = 0u;
        for (int k = 0; k < 8; ++k) st->phase_cycles[k] = 0;
        st->seq_base = a.seq_base;
    }
}

// =============================================================================================
// Generate mode (MPI_OMP.hpp:237-247): local row i is global row g = i + offset;
// A[i][j] = 2 if g == j, 1 if |g - j| == 1, else 0; pad columns [n, lda) are zero.
// Written by the GPU straight into the padded HBM layout, two doubles per store.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) generate_matrix_kernel(void *A_, long long rows, long long n, long long lda, long long offset)
{
    T *A = static_cast<T *>(A_);
    const long long half = lda >> 1;
    const long long total = rows * half;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / half;
        const long long j = (t - i * half) * 2;
        const long long gr = i + offset;
        const long long d0 = gr - j, d1 = gr - (j + 1);
        const T v0 = (j < n) ? (d0 == 0 ? T(2) : ((d0 == 1 || d0 == -1) ? T(1) : T(0))) : T(0);
        const T v1 = (j + 1 < n) ? (d1 == 0 ? T(2) : ((d1 == 1 || d1 == -1) ? T(1) : T(0))) : T(0);
        A[i * lda + j] = v0;
        A[i * lda + j + 1] = v1;
    }
}

// Option matrix_f32: rows of an fp64 source (row pitch src_ld elements) -> the fp32 matrix block (row pitch lda, pad columns
// already zero).  stats[0] counts the entries whose fp32 value differs from the fp64 one (0 = the mixed-storage solve works on
// exactly the caller's matrix), stats[1] the finite entries that overflow to infinity.
__global__ void __launch_bounds__(256) narrow_rows_kernel(const double *__restrict__ src, long long src_ld, float *__restrict__ dst,
                                                          long long lda, long long rows, long long n, unsigned long long *stats)
{
    const long long pairs = (n + 1) >> 1, total = rows * pairs;
    unsigned int inexact = 0, overflow = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / pairs, j = (t - i * pairs) * 2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (j + e >= n) break;
            const double v = src[i * src_ld + j + e];
            const float f = __double2float_rn(v);
            dst[i * lda + j + e] = f;
            if ((double)f != v && v == v) ++inexact; // a NaN stays a NaN: not counted
            if (isinf(f) && !isinf(v)) ++overflow;
        }
    }
    inexact = __reduce_add_sync(0xffffffffu, inexact);
    overflow = __reduce_add_sync(0xffffffffu, overflow);
    if ((threadIdx.x & 31) == 0) {
        if (inexact) atomicAdd(&stats[0], (unsigned long long)inexact);
        if (overflow) atomicAdd(&stats[1], (unsigned long long)overflow);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(void *v_, long long n, long long n_padded, double value)
{
    T *v = static_cast<T *>(v_);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += (long long)gridDim.x * blockDim.x)
        v[i] = i < n ? (T)value : T(0);
}

// Read-only streaming ceiling: sum of every element of the row block with 128-bit loads, in the best access pattern found
// for B200's HBM3e by tools/hbm_read_patterns.cu (profiles/r01_hbm_read_patterns.log: 97 shapes, 3.7-7.44 TB/s): 512 threads,
// 32 independent 16-byte loads per thread in flight, the buffer cut into 4 MB chunks dealt round-robin to 2 x SMs CTAs
// (7439 GB/s; one contiguous segment per CTA, the first version of this kernel, reaches 5.9 TB/s).  This is the control the GEMV is
// compared with in bench.py (`roofline.read_only_stream_GBps`).
constexpr int kStreamThreads = 512, kStreamLoads = 32;
constexpr long long kStreamChunk16 = 4ll * 1024 * 1024 / 16; // 16-byte words per chunk

__global__ void __launch_bounds__(kStreamThreads) stream_read_kernel(const double *A, long long count2, double *partials)
{
    __shared__ double scratch[32];
    const uint64_t pol = l2_policy_evict_first();
    constexpr int NT = kStreamThreads, U = kStreamLoads;
    double s = 0.0;
    for (long long c0 = (long long)blockIdx.x * kStreamChunk16; c0 < count2; c0 += (long long)gridDim.x * kStreamChunk16) {
        const long long c1 = c0 + kStreamChunk16 < count2 ? c0 + kStreamChunk16 : count2;
        long long i = c0 + threadIdx.x;
        for (; i + (long long)(U - 1) * NT < c1; i += (long long)U * NT) {
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_stream_f64x2(A + 2 * (i + (long long)u * NT), pol);
#pragma unroll
            for (int u = 0; u < U; ++u) s += v[u].x + v[u].y;
        }
        for (; i < c1; i += NT) {
            const double2 v = ldg_stream_f64x2(A + 2 * i, pol);
            s += v.x + v.y;
        }
    }
    const double cta = block_sum(s, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = cta;
}

// =============================================================================================
// Persistent single-kernel CG (single rank, n <= 16384): ONE cooperative launch runs the WHOLE solve — no launches, no ramp-up
// and tail per GEMV, scalars never leave the SMs.  Two kernels, chosen by size (lamcg.cu: solve_persistent):
//   * cg_persistent_v4_kernel  (n < 4081)          p in registers, the CTA's rows of A in shared memory, ONE exchange per iteration
//                                                  (all-gather of Ap as tagged words), dots and scalars computed redundantly;
//   * cg_persistent_v3_kernel  (up to n = 16384)   K1's streaming row sweep inside the loop, p in shared memory, two scalar
//                                                  all-reduces + published r per iteration (grid_allgather_sum).
// The first and second generation (round 1: p in shared memory with row tasks; p in registers with two scalar exchanges — 128 k and
// 145 k it/s at n = 2048 against 237 k now) were measured against these and removed; their logs are profiles/r01_persist_gen2.log,
// r01_small_n_*.log and profiles/r02_small_n_gen4*.log.
// Same arithmetic as K1/K2/K3 (unfused multiply-add, fixed summation orders); same reference loop (OMP.hpp:49-91).
// =============================================================================================
struct PersistArgs {
    const double *A;   // [n][lda]
    const double *b;   // [lda] zero padded
    double *x;         // [n] out
    double *r;         // [n] exchange buffer for the r slices
    double *hist;      // nullable
    unsigned long long *ll;      // v3: [2][grid dst][grid src][16] tagged partial words (p.Ap inboxes, then r.r); v4: [2][lda][2] tagged
                                 // entries of the gathered Ap; zeroed by the host
    DevState *st;
    long long n, lda;
    double eps;
    int max_iters, hist_cap;
    int rows_smem;     // v4: the first rows_smem rows of every CTA's block stay resident in shared memory
    int rows_max;      // max rows per CTA (sizes the row-partial arrays)
    int rows_l2keep;   // v3: leading rows of every CTA loaded with the L2 evict-last policy (kept in L2 between iterations)
    int poll_delay;    // v4: cycles between a thread's arrival at the gather and its first poll
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Tagged-word traffic of the persistent loop: strong (gpu-scope) accesses that bypass L1.  Measured: strong
// loads of one thread are NOT pipelined (10 of them in a row cost ~4000 cycles, i.e. ~400 each), so every
// exchange is arranged to need ONE 16-byte load and ONE 16-byte store per thread.
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_v2u64(unsigned long long *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_gpu_v2u64(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// Acquire load: SASS is LDG.STRONG.GPU + CCTL.IVALL (an L1 invalidate), with NO memory barrier — several hundred cycles cheaper than
// a relaxed load followed by __threadfence() (MEMBAR.SC.GPU + ERRBAR + CCTL.IVALL).
__device__ __forceinline__ void ld_acquire_gpu_v2u64(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// Tagged words only ever grow (tag in the upper half, tags increase every iteration, the inbox is zeroed before a solve), so a
// max-reduction delivers them: resolved in L2, and the all-to-all round measured 12 % shorter than with plain strong stores
// (tools/ll_latency.cu: 2526 vs 2858 cycles per round at G = 148).
__device__ __forceinline__ void red_max_gpu_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Polling through the L2 atomic unit ("add 0"): in the G x G all-to-all a round measured 2240 cycles against 2526 with polling
// loads (tools/ll_latency.cu), presumably because the request never touches L1.  Two 8-byte atomics, each half carries its tag.
__device__ __forceinline__ void atom_poll_gpu_2xu64(unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(a) : "l"(p) : "memory");
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(b) : "l"(p + 1) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); } // MEMBAR.ALL.GPU (not .SC)

// Grid-wide "all-gather + fixed-order sum" that doubles as the grid barrier, in ONE L2 round trip.
// Push model: CTA src stores its partial into slot [dst][src] of EVERY CTA's private inbox (thread t
// serves dst = t), as two 8-byte words {tag:32 | half of the double:32} — the NCCL "LL" idea: an
// aligned 8-byte store is single-copy atomic, so a word whose tag matches carries valid data and no
// separate flag or second trip is needed.  Each CTA then polls only its OWN inbox (thread t polls
// source t), so no cache line is polled by more than one SM (a shared flag array polled by all 148
// SMs measured 2x slower than a plain atomic-counter barrier).  The G values go through shared memory
// and are added in index order, so every CTA gets the same bits.  Everything a CTA wrote before the
// call (its r slice) is visible to all CTAs after it.
constexpr int kPersistMaxGrid = 256;
constexpr int kLLStride = 16; // 8-byte words per (dst, src) slot: one 128-byte line each, so a line has one writer and one reader

// kPublishes: the CTA wrote global data (its r slice) that the other CTAs read after this call, so the
// stores need a release fence before and the polls an acquire fence after; the p.Ap exchange moves
// nothing but the tagged words themselves and skips both.
template <bool kPublishes>
__device__ __forceinline__ double grid_allgather_sum(double my_partial_t0, unsigned long long *inbox /* [G dst][G src][2] */,
                                                     unsigned int tag, double *s_gather /* [kPersistMaxGrid] */, double *s_bcast,
                                                     int *err_flag)
{
    const int G = gridDim.x, t = threadIdx.x;
    if (t == 0) *s_bcast = my_partial_t0; // block_sum() left the CTA partial in thread 0 only
    __syncthreads();
    // Stores and polls are issued by DIFFERENT warps (threads 0..G-1 store, threads kPollBase..kPollBase+G-1 poll):
    // a strong load queued behind the same thread's strong store waited ~2500 cycles for it (measured).
    constexpr int kPollBase = 256;
    if (t < G) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(*s_bcast);
        unsigned long long *dst = inbox + ((size_t)t * G + blockIdx.x) * kLLStride;
        // release: ONE acq_rel fence (after the CTA barrier above, so it covers every thread's r stores), then the relaxed deliveries
        if (kPublishes) fence_acq_rel_gpu();
        red_max_gpu_u64(dst, ((unsigned long long)tag << 32) | (bits >> 32));
        red_max_gpu_u64(dst + 1, ((unsigned long long)tag << 32) | (bits & 0xffffffffull));
    }
    if (t >= kPollBase && t < kPollBase + G) {
        const int srcid = t - kPollBase;
        const unsigned long long *src = inbox + ((size_t)blockIdx.x * G + srcid) * kLLStride;
        unsigned long long w0, w1;
        const long long t0 = clock64();
        for (;;) {
            // each 8-byte half carries its own tag, so a torn 16-byte access is harmless; acquire: the polling load itself
            // measured at n = 2048 (profiles/r01_persist_gen2.log): acquire LOADS for the publishing exchange (4820 cycles; 5349 with
            // acquire atomics), relaxed ATOMICS for the other one (4011 cycles; 4324 with relaxed loads)
            if (kPublishes) ld_acquire_gpu_v2u64(src, w0, w1);
            else atom_poll_gpu_2xu64(const_cast<unsigned long long *>(src), w0, w1);
            if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag) break;
            if (clock64() - t0 > 4000000000LL) {
                *err_flag = 3;
                __threadfence_system();
                __trap();
            }
        }
        s_gather[srcid] = __longlong_as_double((long long)(((w0 & 0xffffffffull) << 32) | (w1 & 0xffffffffull)));
    }
    __syncthreads();
    // fixed-order sum by warp 0; the result is returned in ALL LANES OF WARP 0 ONLY (0.0 elsewhere): the caller
    // derives the scalars every thread needs (alpha / beta / stop flag) once, in warp 0, and broadcasts those —
    // fp64 divisions and the square root executed by all 32 warps cost ~1.3 us of FP64 pipe time per iteration.
    double s = 0.0;
    if (t < 32) {
        for (int i = t; i < G; i += 32) s = __dadd_rn(s, s_gather[i]);
        s = warp_sum(s);
    }
    return s;
}

// Measured and dropped (round 1, profiles/r01_persist_gen2.log): a two-level version of this exchange (groups of
// ceil(sqrt(G)) = 13 CTAs all-to-all, then one word per group: 25 messages per CTA instead of 148) was SLOWER, 6.9 k / 8.7 k
// cycles per exchange against 5.0 k / 5.4 k at G = 148.  The cost is per hop (store -> visible in L2 -> seen by a spinning
// strong load -> CTA barrier -> fixed-order sum, ~1800 cycles even with 13 messages; ~700 per gpu-scope fence), not the
// number of messages in flight, so one hop with G messages beats two hops with sqrt(G).

constexpr int kPersistThreads = 512; // 128 registers per thread: the exchange and the GEMV stay spill-free

// =============================================================================================
// Persistent loop, fourth generation ("gathered Ap", n <= 4096; default for n <= 2048).  Generations 1-3 follow the textbook
// distribution of CG: every CTA owns the x, r entries of its rows, so an iteration needs TWO grid-wide scalar all-reduces (p.Ap,
// r.r) and the publication of the new r slices (release fence + acquire polls) — at n = 2048 that was 2 x ~2.3 us of a 6.9 us
// iteration, and neither a two-level exchange through L2 nor a thread-block-cluster first hop over DSMEM
// (tools/cluster_exchange.cu, profiles/r02_cluster_exchange.log) shortens a hop chain that has to cross L2 once anyway.
// This generation exchanges ONCE per iteration: the CTAs all-gather Ap itself.  Each owner thread stores its row's Ap as two
// self-validating 8-byte words {iteration tag : 32 | half of the double : 32} (NCCL's "LL" idea: an aligned 8-byte store is
// single-copy atomic, so a word whose tag matches carries valid data — no flag, no fence, no second trip), every thread of
// every CTA polls the entries of ITS columns (the ones whose p it keeps in registers), and then every CTA computes p.Ap,
// alpha, r -= alpha Ap, r.r, beta, the stop test and p = r + beta p REDUNDANTLY on the full vectors, which live distributed
// over the registers of its 512 threads.  Same operations per element as the reference loop (OMP.hpp:68-78), fixed summation
// orders, identical in every CTA, so all CTAs take the same branch without ever exchanging a scalar.
//   GEMV as in the second generation: warp w owns a column segment of every row of the CTA, p slice in registers, all rows of
//   the CTA resident in shared memory (n = 2048: 14 x 16 KB), 8 rows accumulated at once, halving butterfly.
// =============================================================================================
// Two adjacent tagged entries (columns c, c + 1: 32 bytes) in ONE 256-bit relaxed gpu-scope load (sm_100).  Measured against two
// 128-bit loads and against weak ld.cg polls (profiles/r02_allgather_bench.log, r02_small_n_gen4_final.log): 2235 vs 2431 / 2422
// cycles per gather round, 234 k vs 223 k it/s at n = 2048.
__device__ __forceinline__ void ll_load_pair(const unsigned long long *p, unsigned long long (&w)[4])
{
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p) : "memory");
}

// Sum of one value per thread over the CTA in a fixed order (lane butterfly, then the 16 warps in index order); the total is
// returned in EVERY thread.  s_red: [kPersistThreads / 32] doubles, reusable right after the call (two barriers inside).
__device__ __forceinline__ double persist_block_total(double v, double *s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 512 / 32; ++w) s = __dadd_rn(s, s_red[w]);
    __syncthreads();
    return s;
}

// One group of up to 8 rows of the fourth generation's GEMV over this lane's columns (cbase + 64 k + {0, 1}): acc[j] = partial sum of
// row j.  SMEM: rows in shared memory, else in global memory (L2); FULL: every lane's columns exist.  Straight-line code with
// UNCONDITIONAL loads, so that the RB * K2 loads of a batch are issued back to back (predicated loads were issued one by one: one
// exposed L2 latency per row at n > 2048): rows beyond nrows re-read the group's last row — their sums land in slots of `part` that
// nobody reads — and columns beyond lda re-read a valid column, whose product with this lane's p (zero there) is zero.
template <int PL, bool SMEM, bool FULL>
__device__ __forceinline__ void v4_gemv_group(const double *__restrict__ rowp, int lda, int cbase, int nrows, const double (&preg)[PL], double (&acc)[8])
{
    constexpr int K2 = PL / 2;
    constexpr int RB = PL >= 8 ? 2 : 4; // rows per batch: RB * K2 loads of 16 bytes in flight (8 rows per batch measured slower:
                                        // 3513 vs 3318 cycles for the 14 rows at n = 2048, profiles/r02_gen4_gemv_ab.log)
    if (FULL) lda = 16 * 32 * PL; // FULL means lda == 16 warps x 32 lanes x PL columns: a compile-time row stride, so the loads of a
                                  // full group are addressed by immediate offsets (no per-row multiply / clamp instructions)
    int coff[K2];
#pragma unroll
    for (int k = 0; k < K2; ++k) coff[k] = (FULL || cbase + 64 * k < lda) ? 64 * k : lda - 2 - cbase;
    if (nrows == 8) { // warp-uniform: the common case, no index clamping
#pragma unroll
        for (int h = 0; h < 8; h += RB) {
            double2 av[RB][K2];
#pragma unroll
            for (int j = 0; j < RB; ++j)
#pragma unroll
                for (int k = 0; k < K2; ++k) {
                    const double *src = rowp + (size_t)(h + j) * lda + coff[k];
                    if (SMEM) av[j][k] = *reinterpret_cast<const double2 *>(src);
                    else av[j][k] = __ldg(reinterpret_cast<const double2 *>(src));
                }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < K2; ++k) {
                    s = mul_add(av[j][k].x, preg[2 * k], s);
                    s = mul_add(av[j][k].y, preg[2 * k + 1], s);
                }
                acc[h + j] = s;
            }
        }
        return;
    }
    const int lastrow = nrows - 1;
#pragma unroll
    for (int h = 0; h < 8; h += RB) {
        if (h > 0 && h >= nrows) { // warp-uniform: nothing left in this group
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[h + j] = 0.0;
            continue;
        }
        double2 av[RB][K2];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const double *rp = rowp + (size_t)(h + j < lastrow ? h + j : lastrow) * lda;
#pragma unroll
            for (int k = 0; k < K2; ++k) {
                if (SMEM) av[j][k] = *reinterpret_cast<const double2 *>(rp + coff[k]);
                else av[j][k] = __ldg(reinterpret_cast<const double2 *>(rp + coff[k]));
            }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K2; ++k) {
                s = mul_add(av[j][k].x, preg[2 * k], s);
                s = mul_add(av[j][k].y, preg[2 * k + 1], s);
            }
            acc[h + j] = s;
        }
    }
}

template <int PL>
__global__ void __launch_bounds__(512, 1) cg_persistent_v4_kernel(PersistArgs a)
{
    extern __shared__ __align__(16) double psm[];
    constexpr int NT = 512, NW = NT / 32;
    constexpr int K2 = PL / 2;   // column pairs (16-byte loads of A, 32-byte tagged entries of Ap) per lane
    constexpr int SEG = 32 * PL; // columns per warp
    constexpr unsigned FULL = 0xffffffffu;
    const int n = (int)a.n, lda = (int)a.lda;
    double *arows = psm;                              // [rows_smem][lda] resident rows of A
    double *part = psm + (size_t)a.rows_smem * lda;   // [NW][rows_max rounded up to 8] per-warp row partials
    __shared__ double s_red[NW];
    __shared__ double s_bcast;
    __shared__ double s_scal[4]; // alpha | beta | rr | stop code

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int base = n / G, rem = n % G;
    const int r0 = bid * base + (bid < rem ? bid : rem);
    const int rcnt = base + (bid < rem ? 1 : 0);
    const int cbase = warp * SEG + 2 * lane; // this lane's columns: cbase + 64 k + {0, 1}
    const int rows_pad = (a.rows_max + 7) & ~7;
    const bool full_cols = lda == NW * SEG; // every lane's columns lie inside the (padded) matrix
    DevState *st = a.st;

    // ---- init: p = r = b in registers (b is zero padded to lda), own x = 0, bb = b.b (same order in every CTA)
    double preg[PL], rreg[PL];
    double local = 0.0;
#pragma unroll
    for (int k = 0; k < K2; ++k) {
        const int c = cbase + 64 * k;
        double2 bv = make_double2(0.0, 0.0);
        if (c < lda) bv = *reinterpret_cast<const double2 *>(a.b + c);
        preg[2 * k] = rreg[2 * k] = bv.x;
        preg[2 * k + 1] = rreg[2 * k + 1] = bv.y;
        local = mul_add(bv.x, bv.x, local);
        local = mul_add(bv.y, bv.y, local);
    }
    const double bb = persist_block_total(local, s_red);
    double x_own = 0.0, r_own = 0.0, p_own = 0.0, Ap_own = 0.0; // the owner thread's copies of its row's entries
    if (tid < rcnt) r_own = p_own = a.b[r0 + tid];
    const int nres = rcnt < a.rows_smem ? rcnt : a.rows_smem;
    {
        const double *src = a.A + (size_t)r0 * lda;
        for (int i = 2 * tid; i < nres * lda; i += 2 * NT)
            *reinterpret_cast<double2 *>(arows + i) = __ldg(reinterpret_cast<const double2 *>(src + i));
    }
    __syncthreads();

    double rr = bb, beta = 0.0;
    int it;
    int poll_delay = a.poll_delay; // cycles between arriving at the gather and the first poll (grows when that poll comes too early)
    bool converged = false, broke = false;
    long long ph[6] = {0, 0, 0, 0, 0, 0}; // p update | GEMV | row sums + publish | gather Ap | p.Ap, alpha, r, r.r | beta, stop test
    long long tc = clock64();
#define LAMCG_PHASE(k) { const long long now_ = clock64(); ph[k] += now_ - tc; tc = now_; }
    for (it = 1; it <= a.max_iters; ++it) {
        if (it > 1) { // p = r + beta p on the register slices (axpby(1.0, r, beta, p), OMP.hpp:78)
#pragma unroll
            for (int k = 0; k < PL; ++k) preg[k] = __dadd_rn(rreg[k], __dmul_rn(beta, preg[k]));
            if (tid < rcnt) p_own = __dadd_rn(r_own, __dmul_rn(beta, p_own));
        }
        LAMCG_PHASE(0)
        // ---- GEMV: 8 rows at a time over this warp's column segment
        for (int g0 = 0; g0 < rcnt; g0 += 8) {
            double acc[8];
            const int nrows = rcnt - g0 < 8 ? rcnt - g0 : 8;
            const bool all_smem = g0 + nrows <= nres, all_global = g0 >= nres;
            if (all_smem || all_global) {
                // fast paths (every row of the group comes from the same place: shared memory — the n = 2048 layout — or L2):
                // straight-line code, RB rows x K2 16-byte loads issued before their products; row validity is a warp-uniform
                // predicate, column validity (only when the matrix is narrower than the 16 warps' segments) a per-lane one.  The
                // branchy general path below exposed one load latency per row (n = 2048: 3.9 k cycles for 14 rows, 3.35 k here).
                if (all_smem) {
                    if (full_cols) v4_gemv_group<PL, true, true>(arows + (size_t)g0 * lda + cbase, lda, cbase, nrows, preg, acc);
                    else v4_gemv_group<PL, true, false>(arows + (size_t)g0 * lda + cbase, lda, cbase, nrows, preg, acc);
                } else {
                    if (full_cols) v4_gemv_group<PL, false, true>(a.A + (size_t)(r0 + g0) * lda + cbase, lda, cbase, nrows, preg, acc);
                    else v4_gemv_group<PL, false, false>(a.A + (size_t)(r0 + g0) * lda + cbase, lda, cbase, nrows, preg, acc);
                }
            } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j] = 0.0;
                const int row = g0 + j;
                if (row < nres) {
                    const double *arow = arows + row * lda;
#pragma unroll
                    for (int k = 0; k < K2; ++k) {
                        const int c = cbase + 64 * k;
                        if (c < lda) {
                            const double2 av = *reinterpret_cast<const double2 *>(arow + c);
                            acc[j] = mul_add(av.x, preg[2 * k], acc[j]);
                            acc[j] = mul_add(av.y, preg[2 * k + 1], acc[j]);
                        }
                    }
                } else if (row < rcnt) {
                    const double *arow = a.A + (size_t)(r0 + row) * lda;
#pragma unroll
                    for (int k = 0; k < K2; ++k) {
                        const int c = cbase + 64 * k;
                        if (c < lda) {
                            const double2 av = __ldg(reinterpret_cast<const double2 *>(arow + c));
                            acc[j] = mul_add(av.x, preg[2 * k], acc[j]);
                            acc[j] = mul_add(av.y, preg[2 * k + 1], acc[j]);
                        }
                    }
                }
            }
            }
            // halving butterfly: each kept value is computed by exactly one lane (see the second generation)
            const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
            double v4[4], v2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double keep = b4 ? acc[4 + i] : acc[i], send = b4 ? acc[i] : acc[4 + i];
                v4[i] = __dadd_rn(keep, __shfl_xor_sync(FULL, send, 16));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double keep = b3 ? v4[2 + i] : v4[i], send = b3 ? v4[i] : v4[2 + i];
                v2[i] = __dadd_rn(keep, __shfl_xor_sync(FULL, send, 8));
            }
            double v = __dadd_rn(b2 ? v2[1] : v2[0], __shfl_xor_sync(FULL, b2 ? v2[0] : v2[1], 4));
            v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 2));
            v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 1));
            if ((lane & 3) == 0) part[warp * rows_pad + g0 + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = v;
        }
        __syncthreads();
        LAMCG_PHASE(1)
        // ---- the owner of a row adds its 16 warp partials in warp order and publishes the row's Ap as two tagged words
        // layout [iteration parity][lda][2 words]: double-buffered by parity (WAR: see below)
        unsigned long long *ll = a.ll + (size_t)(it & 1) * 2 * (size_t)lda;
        const unsigned long long tag = (unsigned long long)(unsigned int)it << 32;
        if (tid < rcnt) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) sum = __dadd_rn(sum, part[w * rows_pad + tid]);
            Ap_own = sum;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(sum);
            st_relaxed_gpu_v2u64(ll + 2 * (size_t)(r0 + tid), tag | (bits >> 32), tag | (bits & 0xffffffffull));
        }
        LAMCG_PHASE(2)
        // ---- all-gather: poll the entries of this thread's columns.  A CTA can run at most one iteration ahead of the slowest
        // one (it cannot finish iteration it+1 without that CTA's words of it+1, which are stored after its polls of iteration it
        // have completed), so two buffers are enough and a matching tag always belongs to the current iteration.
        double ap[PL];
        {
            unsigned long long w[K2][4];
            bool ok[K2];
#pragma unroll
            for (int k = 0; k < K2; ++k) ok[k] = cbase + 64 * k >= n; // nothing beyond n (n is even or the odd tail is handled below)
            const long long t0 = clock64();
            // The words need one L2 hop (~800 cycles) after the slowest CTA's store.  Polling earlier only puts 148 x 512 load
            // requests in front of those stores and costs a whole extra round (n = 1024: 254 k it/s with a first poll after 400
            // cycles, 349 k after 500; profiles/r02_gen4_delay_sweep.log), so a thread waits before its first poll, and waits
            // longer from then on whenever that first poll still found a stale entry (slower part, other clocks).  Measured and
            // dropped (profiles/r02_gen4_probe.log, r02_gen4_poll_sweep.log): 2-8 replicas of the gathered array (they only helped
            // while the polls started too early: 180 -> 218 k it/s at n = 2048, against 237 k with the delay and one copy), stores
            // staged through shared memory, __nanosleep back-off between rounds (168-180 k).
            if (poll_delay > 0)
                while (clock64() - t0 < poll_delay) {}
            bool first_round = true;
            for (;;) {
                bool all = true;
#pragma unroll
                for (int k = 0; k < K2; ++k)
                    if (!ok[k]) ll_load_pair(ll + 2 * (size_t)(cbase + 64 * k), w[k]);
#pragma unroll
                for (int k = 0; k < K2; ++k)
                    if (!ok[k]) {
                        const int c = cbase + 64 * k;
                        const bool second = c + 1 < n; // the last column pair of an odd n has only one entry
                        ok[k] = (w[k][0] >> 32) == (tag >> 32) && (w[k][1] >> 32) == (tag >> 32) &&
                                (!second || ((w[k][2] >> 32) == (tag >> 32) && (w[k][3] >> 32) == (tag >> 32)));
                        all = all && ok[k];
                    }
                if (all) break;
                if (first_round && a.poll_delay > 0 && poll_delay < 4000) poll_delay += 64;
                first_round = false;
                if (clock64() - t0 > 4000000000LL) {
                    st->error = 3;
                    __threadfence_system();
                    __trap();
                }
            }
#pragma unroll
            for (int k = 0; k < K2; ++k) {
                const int c = cbase + 64 * k;
                ap[2 * k] = c < n ? __longlong_as_double((long long)(((w[k][0] & 0xffffffffull) << 32) | (w[k][1] & 0xffffffffull))) : 0.0;
                ap[2 * k + 1] = c + 1 < n ? __longlong_as_double((long long)(((w[k][2] & 0xffffffffull) << 32) | (w[k][3] & 0xffffffffull))) : 0.0;
            }
        }
        LAMCG_PHASE(3)
        // ---- redundant vector work on the register slices: p.Ap, alpha, r -= alpha Ap, r.r
        local = 0.0;
#pragma unroll
        for (int k = 0; k < PL; ++k) local = mul_add(preg[k], ap[k], local);
        // block total in the fixed order (lane butterfly, then the 16 warps in index order), finished by warp 0 alone, which
        // also derives the scalar: two CTA barriers per reduction instead of three
        local = warp_sum(local);
        if (lane == 0) s_red[warp] = local;
        __syncthreads();
        if (warp == 0) {
            double pAp = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) pAp = __dadd_rn(pAp, s_red[w]);
            if (lane == 0) s_scal[0] = rr / pAp; // alpha = rr / (p.Ap): one division per CTA, broadcast
        }
        __syncthreads();
        const double alpha = s_scal[0], nalpha = -alpha;
        local = 0.0;
#pragma unroll
        for (int k = 0; k < PL; ++k) {
            rreg[k] = __dadd_rn(__dmul_rn(nalpha, ap[k]), rreg[k]); // axpby(-alpha, Ap, 1.0, r)
            local = mul_add(rreg[k], rreg[k], local);
        }
        if (tid < rcnt) {
            x_own = __dadd_rn(__dmul_rn(alpha, p_own), x_own);       // axpby(alpha, p, 1.0, x)
            r_own = __dadd_rn(__dmul_rn(nalpha, Ap_own), r_own);
        }
        local = warp_sum(local);
        if (lane == 0) s_red[warp] = local; // (every thread has read alpha past the barrier above; s_red's readers were warp 0 before it)
        __syncthreads();
        LAMCG_PHASE(4)
        if (warp == 0) {
            double rrn = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) rrn = __dadd_rn(rrn, s_red[w]);
            if (lane == 0) {
                const double rel0 = sqrt(rrn / bb);
                s_scal[1] = rrn / rr; // beta = rr_new / rr
                s_scal[2] = rrn;
                const bool broke0 = !(rel0 == rel0) || isinf(rel0) || !(s_scal[1] == s_scal[1]);
                s_scal[3] = rel0 < a.eps ? 1.0 : (broke0 ? 2.0 : 0.0);
                if (bid == 0 && a.hist && it - 1 < a.hist_cap) a.hist[it - 1] = rel0;
            }
        }
        __syncthreads();
        beta = s_scal[1];
        rr = s_scal[2];
        const double code = s_scal[3];
        LAMCG_PHASE(5)
        if (code == 1.0) { converged = true; break; }
        if (code == 2.0) { broke = true; break; }
        // no barrier needed here: thread 0 rewrites s_scal only after the CTA barriers of the next GEMV / block sum
    }
    if (tid < rcnt) a.x[r0 + tid] = x_own;
    if (bid == 0 && tid == 0) {
        st->bb = bb;
        st->rr_final = rr;
        st->iters_done = (converged || broke) ? it : (a.max_iters > 0 ? a.max_iters : 0);
        st->converged = converged ? 1 : 0;
        st->breakdown = broke ? 1 : 0;
        st->max_iters = a.max_iters;
        st->eps = a.eps;
        st->done = 1;
        for (int k = 0; k < 6; ++k) st->phase_cycles[k] = ph[k];
    }
#undef LAMCG_PHASE
}

// =============================================================================================
// Persistent loop, third generation ("streaming", 2048 < n <= 16384): the matrix no longer fits in the shared memory of
// the 148 SMs, so the GEMV inside the one-kernel loop is K1's CTA-wide row sweep (R rows in flight, U 16-byte streaming
// loads per row, thread and chunk; p comes from shared memory once per chunk and is reused for the R rows), fed from
// L2 / HBM.  Against the graph loop this saves the three launches per iteration and K1's ramp-up and tail at sizes
// where a GEMV lasts 20-120 us; against the first generation (one warp per row task) it streams at K1's rate.
// Everything else (p in shared memory, two tagged-word all-gathers, redundant scalars, fixed summation orders) is the
// first generation's.
// =============================================================================================
template <int R, int U>
__global__ void __launch_bounds__(kPersistThreads, 1) cg_persistent_v3_kernel(PersistArgs a)
{
    extern __shared__ __align__(16) double psm[];
    constexpr int NT = kPersistThreads, NWARP = NT / 32;
    constexpr int CH = NT * U * 2; // columns per chunk
    static_assert(R <= 32, "row sums are finished by one warp");
    const int n = (int)a.n, lda = (int)a.lda;
    double *p = psm;            // [lda]
    double *Ap_s = psm + lda;   // [rows_max rounded up to R]
    __shared__ double red[NWARP][R];
    __shared__ double scratch[32];
    __shared__ double s_bcast;
    __shared__ double s_gather[kPersistMaxGrid];
    __shared__ double s_scal[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int base = n / G, rem = n % G;
    const int r0 = bid * base + (bid < rem ? bid : rem);
    const int rcnt = base + (bid < rem ? 1 : 0);
    DevState *st = a.st;
    // a matrix that fits in L2 should stay there between iterations; a larger one is streamed (evict-first) so that it does
    // not push r and the exchange words out
    // a matrix that fits in L2 stays there between iterations; of a larger one the first rows_l2keep rows of every CTA are loaded
    // evict-last (together ~80 MB of the 126 MB L2: that slice is served from L2 in every iteration after the first) and the rest is
    // streamed evict-first so that it does not push that slice, r and the exchange words out
    const bool fits_l2 = (size_t)n * lda * sizeof(double) <= (size_t)96 << 20;
    const uint64_t polStream = fits_l2 ? l2_policy_evict_normal() : l2_policy_evict_first();
    const uint64_t polKeep = fits_l2 ? l2_policy_evict_normal() : l2_policy_evict_last();

    double local = 0.0;
    for (int i = tid; i < lda; i += NT) {
        const double bi = a.b[i];
        p[i] = bi;
        local = mul_add(bi, bi, local);
    }
    const double bb_t0 = block_sum(local, scratch);
    if (tid == 0) s_bcast = bb_t0;
    __syncthreads();
    const double bb = s_bcast;
    double x_own = 0.0, r_own = 0.0, Ap_own = 0.0;
    if (tid < rcnt) r_own = a.b[r0 + tid];

    double rr = bb, beta = 0.0;
    int it;
    bool converged = false, broke = false;
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    long long tc = clock64();
#define LAMCG_PHASE(k) { const long long now_ = clock64(); ph[k] += now_ - tc; tc = now_; }
    for (it = 1; it <= a.max_iters; ++it) {
        if (it > 1) {
            for (int i = tid; i < n; i += NT) p[i] = __dadd_rn(__ldcg(&a.r[i]), __dmul_rn(beta, p[i]));
            __syncthreads();
        }
        LAMCG_PHASE(0)
        for (int pr = 0; pr < rcnt; pr += R) {
            const int nr = rcnt - pr < R ? rcnt - pr : R;
            const double *arow0 = a.A + (size_t)(r0 + pr) * lda;
            double acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0.0;
            for (int c0 = 0; c0 < lda; c0 += CH) {
                double2 pv[U];
                bool cv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + 2 * tid + u * NT * 2;
                    cv[u] = c < lda;
                    pv[u] = cv[u] ? *reinterpret_cast<const double2 *>(p + c) : make_double2(0.0, 0.0);
                }
                double2 av[R][U];
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int c = c0 + 2 * tid + u * NT * 2;
                        av[r][u] = (cv[u] && r < nr) ? ldg_stream_f64x2(arow0 + (size_t)r * lda + c, pr + r < a.rows_l2keep ? polKeep : polStream) : make_double2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        acc[r] = mul_add(av[r][u].x, pv[u].x, acc[r]);
                        acc[r] = mul_add(av[r][u].y, pv[u].y, acc[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) red[warp][r] = acc[r];
            }
            __syncthreads();
            if (warp == 0 && lane < nr) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) sum = __dadd_rn(sum, red[w][lane]);
                Ap_s[pr + lane] = sum;
            }
            __syncthreads();
        }
        LAMCG_PHASE(1)
        double contrib = 0.0;
        if (tid < rcnt) {
            Ap_own = Ap_s[tid];
            contrib = __dmul_rn(p[r0 + tid], Ap_own);
        }
        {
            const double cta_pap = a.rows_max <= 32 ? (warp == 0 ? warp_sum(contrib) : 0.0) : block_sum(contrib, scratch);
            const double pAp_w0 = grid_allgather_sum<false>(cta_pap, a.ll, (unsigned int)it, s_gather, &s_bcast, &st->error);
            if (tid == 0) s_scal[0] = rr / pAp_w0;
        }
        LAMCG_PHASE(2)
        __syncthreads();
        const double alpha = s_scal[0];
        LAMCG_PHASE(3)
        contrib = 0.0;
        if (tid < rcnt) {
            x_own = __dadd_rn(__dmul_rn(alpha, p[r0 + tid]), x_own);
            r_own = __dadd_rn(__dmul_rn(-alpha, Ap_own), r_own);
            __stcg(&a.r[r0 + tid], r_own);
            contrib = __dmul_rn(r_own, r_own);
        }
        const double cta_rr = a.rows_max <= 32 ? (warp == 0 ? warp_sum(contrib) : 0.0) : block_sum(contrib, scratch);
        const double rrn_w0 = grid_allgather_sum<true>(cta_rr, a.ll + (size_t)kLLStride * G * G, (unsigned int)it, s_gather, &s_bcast, &st->error);
        LAMCG_PHASE(4)
        if (tid == 0) {
            const double rel0 = sqrt(rrn_w0 / bb);
            s_scal[1] = rrn_w0 / rr;
            s_scal[2] = rrn_w0;
            const bool broke0 = !(rel0 == rel0) || isinf(rel0) || !(s_scal[1] == s_scal[1]);
            s_scal[3] = rel0 < a.eps ? 1.0 : (broke0 ? 2.0 : 0.0);
            if (bid == 0 && a.hist && it - 1 < a.hist_cap) a.hist[it - 1] = rel0;
        }
        __syncthreads();
        beta = s_scal[1];
        rr = s_scal[2];
        LAMCG_PHASE(5)
        if (s_scal[3] == 1.0) { converged = true; break; }
        if (s_scal[3] == 2.0) { broke = true; break; }
    }
    if (tid < rcnt) a.x[r0 + tid] = x_own;
    if (bid == 0 && tid == 0) {
        st->bb = bb;
        st->rr_final = rr;
        st->iters_done = (converged || broke) ? it : (a.max_iters > 0 ? a.max_iters : 0);
        st->converged = converged ? 1 : 0;
        st->breakdown = broke ? 1 : 0;
        st->max_iters = a.max_iters;
        st->eps = a.eps;
        st->done = 1;
        for (int k = 0; k < 6; ++k) st->phase_cycles[k] = ph[k];
    }
#undef LAMCG_PHASE
}

} // namespace lamcgk
