// cluster_exchange.cu — microbenchmark for a thread-block-cluster first hop in the persistent loop's scalar exchange
// (measurement tool, not product code; companion of ll_latency.cu).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_exchange tools/cluster_exchange.cu && /tmp/cluster_exchange
// The flat exchange (grid_allgather_sum): every CTA pushes a tagged word to all G inboxes through L2, G x G messages, one hop
// (~2240 cycles per round at G = 148, ll_latency.cu).  Here, for cluster sizes C = 1, 2, 4, 8:
//   hop 1 (DSMEM): every CTA of a cluster sends its value to the cluster leader's shared memory with st.async + the leader's
//                  mbarrier (data and signal in one instruction, no cluster-wide barrier);
//   variant B:     the leader pushes the cluster sum to ALL G inboxes through L2 (NC x G messages), every CTA polls NC slots;
//   variant A:     leaders exchange among themselves through L2 (NC x NC), then send the total back to their members over DSMEM.
// Also: round trip of barrier.cluster arrive+wait, and of one st.async ping-pong inside a cluster.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;
constexpr int kStride = 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(u64 *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity)
{
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity))
        if (clock64() - t0 > 400000000LL) { printf("mbar timeout block %d\n", blockIdx.x); __trap(); }
}
// 8 bytes into a peer CTA's shared memory, completion counted on that CTA's mbarrier
__device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void red_max(u64 *p, u64 v) { asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void atom_poll(u64 *p, u64 &a, u64 &b)
{
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(a) : "l"(p) : "memory");
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(b) : "l"(p + 1) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- (1) cluster barrier round
__global__ void __launch_bounds__(512, 1) k_cluster_barrier(int rounds, long long *cycles)
{
    const long long t0 = clock64();
    for (int i = 0; i < rounds; ++i) { cluster_arrive(); cluster_wait(); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

// ---- (2) hierarchical exchange.  VARIANT 0 = B (leader -> all CTAs through L2), 1 = A (leaders all-to-all + DSMEM broadcast)
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k_exchange(u64 *inbox, int rounds, long long *cycles, double *out)
{
    __shared__ __align__(16) double slots[2][16]; // leader: the C values of its cluster, double-buffered by round parity (a member
                                                  // may send round i+1 while slow leader threads still read round i)
    __shared__ __align__(8) u64 bar_gather;       // leader: completes when C values have landed
    __shared__ __align__(16) double total_slot[2]; // member: the grid total from its leader (variant A)
    __shared__ __align__(8) u64 bar_total;
    __shared__ double s_gather[160];
    const int t = threadIdx.x, G = gridDim.x;
    const unsigned C = cluster_nctarank(), crank = cluster_ctarank();
    const int NC = G / (int)C, cid = blockIdx.x / (int)C;
    if (t == 0) { mbar_init(&bar_gather, 1); mbar_init(&bar_total, 1); }
    __syncthreads();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cluster_arrive(); cluster_wait();   // everybody's barriers exist before the first remote store
    const unsigned leader_slots = mapa(smem_u32(slots), 0), leader_bar = mapa(smem_u32(&bar_gather), 0);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        const unsigned par = (unsigned)(i - 1) & 1u;
        const double mine = 1.0 + blockIdx.x * 1e-3 + i; // stands for the CTA partial
        // hop 1: value -> leader's shared memory
        if (t == 0) {
            if (crank == 0) mbar_expect_tx(&bar_gather, 8 * C);
            st_async_f64(leader_slots + 8 * (16 * par + crank), mine, leader_bar);
        }
        const u64 tag = (u64)i << 32;
        if (crank == 0) {
            // hop 2: every pushing thread of the leader waits for the C values itself and adds them in rank order
            const int ndst = VARIANT == 0 ? G : NC;
            if (t < ndst) {
                mbar_wait(&bar_gather, par);
                double s = 0.0;
                for (unsigned c = 0; c < C; ++c) s += slots[par][c];
                const u64 bits = (u64)__double_as_longlong(s);
                const int dst_cta = VARIANT == 0 ? t : t * (int)C; // all CTAs | leaders only
                u64 *dst = inbox + ((size_t)dst_cta * NC + cid) * kStride;
                red_max(dst, tag | (bits >> 32));
                red_max(dst + 1, tag | (bits & 0xffffffffull));
            }
        }
        double total = 0.0;
        if (VARIANT == 0 || crank == 0) {
            if (t >= 256 && t < 256 + NC) {
                u64 *src = inbox + ((size_t)blockIdx.x * NC + (t - 256)) * kStride;
                u64 w0, w1;
                const long long w = clock64();
                do { atom_poll(src, w0, w1); } while (((w0 >> 32) < (u64)i || (w1 >> 32) < (u64)i) && clock64() - w < 400000000LL);
                s_gather[t - 256] = __longlong_as_double((long long)(((w0 & 0xffffffffull) << 32) | (w1 & 0xffffffffull)));
            }
            __syncthreads();
            if (t < 32) {
                double s = 0.0;
                for (int k = t; k < NC; k += 32) s += s_gather[k];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                total = s;
            }
        }
        if (VARIANT == 1) {
            // hop 3: leader -> members over DSMEM
            if (t == 0) mbar_expect_tx(&bar_total, 8);
            if (crank == 0 && t < (int)C) st_async_f64(mapa(smem_u32(&total_slot[par]), t), total, mapa(smem_u32(&bar_total), t));
            mbar_wait(&bar_total, par);
            total = total_slot[par];
        }
        acc += total;
        __syncthreads();
    }
    if (t == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
    if (t == 0) out[blockIdx.x] = acc;
    cluster_arrive(); cluster_wait(); // no CTA exits while a peer may still write into its shared memory
}

template <typename K>
bool launch(K kernel, int grid, int C, void **args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(512);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)kernel, args);
    if (e != cudaSuccess) { printf("  launch (grid %d, cluster %d) failed: %s\n", grid, C, cudaGetErrorString(e)); cudaGetLastError(); return false; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel (grid %d, cluster %d) failed: %s\n", grid, C, cudaGetErrorString(e)); exit(1); }
    return true;
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int SMs = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, SMs);
    u64 *inbox;
    long long *d_cyc;
    double *d_out;
    const size_t bytes = (size_t)SMs * SMs * kStride * sizeof(u64);
    CK(cudaMalloc(&inbox, bytes));
    CK(cudaMalloc(&d_cyc, sizeof(long long)));
    CK(cudaMalloc(&d_out, SMs * sizeof(double)));
    CK(cudaFuncSetAttribute((const void *)k_exchange<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute((const void *)k_exchange<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute((const void *)k_cluster_barrier, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int rounds = 2000;
    for (int C : {1, 2, 4, 8, 16}) {
        // how many clusters of this size can be resident at once (1 CTA of 512 threads per SM)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(SMs / C * C);
        cfg.blockDim = dim3(512);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int maxc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, (const void *)k_exchange<0>, &cfg);
        if (e != cudaSuccess) { printf("cluster %2d: cudaOccupancyMaxActiveClusters: %s\n", C, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        const int grid = std::min(maxc, SMs / C) * C;
        printf("cluster %2d: %3d clusters co-resident -> grid %3d CTAs", C, maxc, grid);
        if (grid == 0) { printf("\n"); continue; }
        long long cyc;
        void *a0[] = {&rounds, &d_cyc};
        if (launch(k_cluster_barrier, grid, C, a0)) {
            CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
            printf("  | barrier.cluster round %5.0f cyc", (double)cyc / rounds);
        }
        for (int v = 0; v < 2; ++v) {
            CK(cudaMemset(inbox, 0, bytes));
            void *a1[] = {&inbox, &rounds, &d_cyc, &d_out};
            const bool ok = v == 0 ? launch(k_exchange<0>, grid, C, a1) : launch(k_exchange<1>, grid, C, a1);
            if (!ok) continue;
            CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
            double o0, o1;
            CK(cudaMemcpy(&o0, d_out, sizeof o0, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&o1, d_out + grid - 1, sizeof o1, cudaMemcpyDeviceToHost));
            printf("  | %s %5.0f cyc/round%s", v == 0 ? "B: leader->all via L2" : "A: leaders a2a + DSMEM bcast", (double)cyc / rounds, o0 == o1 ? "" : " (MISMATCH)");
        }
        printf("\n");
    }
    return 0;
}
