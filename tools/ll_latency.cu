// ll_latency.cu — microbenchmark behind the persistent loop's scalar exchange (measurement tool, not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ll_latency tools/ll_latency.cu && /tmp/ll_latency
// (1) ping-pong between two CTAs: one-way latency of "store a tagged word -> a spinning load on another SM sees it"
//     for several store / load flavours;
// (2) the all-to-all of grid_allgather_sum stripped to its memory traffic: G CTAs, thread t < G stores one 16-byte
//     tagged word into CTA t's inbox, thread 256 + s polls source s; cycles per round for the same flavours.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef unsigned long long u64;
constexpr int kStride = 16; // words per slot: one 128-byte line each

template <int ST>
__device__ __forceinline__ void put(u64 *p, u64 a, u64 b)
{
    if (ST == 0) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (ST == 1) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory"); __threadfence(); }
    if (ST == 2) { asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
                   asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p + 1), "l"(b) : "memory"); }
    if (ST == 3) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (ST == 4) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (ST == 5) { u64 o; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(a) : "memory");
                   asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p + 1), "l"(b) : "memory"); }
}
template <int LD>
__device__ __forceinline__ void get(const u64 *p, u64 &a, u64 &b)
{
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    if (LD == 1) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    if (LD == 3) { asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(a) : "l"(p) : "memory");
                   asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(b) : "l"(p + 1) : "memory"); }
    if (LD == 4) asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ unsigned smid()
{
    unsigned r;
    asm("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

// ---- (1) ping-pong: CTA `a` and CTA `b` bounce a tag; every other CTA idles.
template <int ST, int LD>
__global__ void pingpong(u64 *slots, int a, int b, int rounds, long long *cycles, unsigned *sm)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != a && me != b) return;
    u64 *mine = slots + (size_t)me * kStride, *theirs = slots + (size_t)(me == a ? b : a) * kStride;
    sm[me == a ? 0 : 1] = smid();
    u64 w0, w1;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        const u64 tag = (u64)i << 32;
        if (me == a) {
            put<ST>(theirs, tag | 1, tag | 2);
            const long long w = clock64();
            do { get<LD>(mine, w0, w1); } while (((w0 >> 32) != (u64)i || (w1 >> 32) != (u64)i) && clock64() - w < 200000000LL);
        } else {
            const long long w = clock64();
            do { get<LD>(mine, w0, w1); } while (((w0 >> 32) != (u64)i || (w1 >> 32) != (u64)i) && clock64() - w < 200000000LL);
            put<ST>(theirs, tag | 1, tag | 2);
        }
    }
    if (me == a) *cycles = clock64() - t0;
}

// ---- (2) all-to-all rounds, layout and thread roles of grid_allgather_sum
template <int ST, int LD>
__global__ void __launch_bounds__(512, 1) alltoall(u64 *inbox, int rounds, long long *cycles)
{
    const int G = gridDim.x, t = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        const u64 tag = (u64)i << 32;
        if (t < G) put<ST>(inbox + ((size_t)t * G + blockIdx.x) * kStride, tag | 1, tag | 2);
        if (t >= 256 && t < 256 + G) {
            const u64 *src = inbox + ((size_t)blockIdx.x * G + (t - 256)) * kStride;
            u64 w0, w1;
            // a fast CTA may already have stored round i + 1 here (the solver alternates two inboxes, which rules that out): accept >= i
            const long long w = clock64();
            do { get<LD>(src, w0, w1); } while (((w0 >> 32) < (u64)i || (w1 >> 32) < (u64)i) && clock64() - w < 200000000LL);
        }
        __syncthreads();
    }
    if (t == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

template <int ST, int LD>
void run(const char *name, u64 *buf, size_t bytes, long long *d_cyc, unsigned *d_sm, int G)
{
    const int rounds = 2000;
    long long cyc;
    unsigned sm[2];
    printf("%-44s", name);
    const int partners[4] = {1, 2, G / 2, G - 1};
    for (int k = 0; k < 4; ++k) {
        CK(cudaMemset(buf, 0, bytes));
        pingpong<ST, LD><<<G, 32>>>(buf, 0, partners[k], rounds, d_cyc, d_sm);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(sm, d_sm, sizeof sm, cudaMemcpyDeviceToHost));
        printf("  sm%3u<->sm%3u %6.0f", sm[0], sm[1], (double)cyc / rounds / 2);
    }
    CK(cudaMemset(buf, 0, bytes));
    void *args[] = {&buf, (void *)&rounds, &d_cyc};
    CK(cudaLaunchCooperativeKernel((void *)alltoall<ST, LD>, dim3(G), dim3(512), args, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    printf("  | all-to-all G=%d: %6.0f cycles/round\n", G, (double)cyc / rounds);
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    const size_t bytes = (size_t)G * G * kStride * sizeof(u64);
    u64 *buf;
    long long *d_cyc;
    unsigned *d_sm;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&d_cyc, sizeof(long long)));
    CK(cudaMalloc(&d_sm, 2 * sizeof(unsigned)));
    printf("%s, %d SMs; one-way hop latency in SM cycles (ping-pong, half a round trip) and all-to-all round time\n", prop.name, G);
    run<0, 0>("st.relaxed.gpu       / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<1, 0>("st.relaxed.gpu+fence / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<2, 0>("red.max.u64 x2       / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<3, 0>("st.volatile          / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<4, 0>("st.cg                / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<5, 0>("atom.exch x2         / ld.relaxed.gpu", buf, bytes, d_cyc, d_sm, G);
    run<0, 1>("st.relaxed.gpu       / ld.volatile", buf, bytes, d_cyc, d_sm, G);
    // ld.global.cv never left its 200 M-cycle watchdog on B200 (the line stays in L1): not a usable polling load
    run<0, 3>("st.relaxed.gpu       / atom.add 0 x2", buf, bytes, d_cyc, d_sm, G);
    run<0, 4>("st.relaxed.gpu       / ld.acquire.gpu", buf, bytes, d_cyc, d_sm, G);
    run<2, 3>("red.max.u64 x2       / atom.add 0 x2", buf, bytes, d_cyc, d_sm, G);
    return 0;
}
