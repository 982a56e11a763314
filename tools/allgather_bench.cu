// allgather_bench.cu — microbenchmark for the fourth-generation persistent CG loop (measurement tool, not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/allgather_bench.out tools/allgather_bench.cu && tools/allgather_bench.out
// Question: how long does a grid-wide ALL-GATHER of an n-vector take inside one cooperative kernel when every value travels as
// two self-validating 8-byte words {tag:32 | half of the double:32} (NCCL "LL"), so that no flag, fence or second trip is needed?
// If it costs about one scalar exchange (~2200-2900 cycles at G = 148, ll_latency.cu), gathering Ap and computing p.Ap, r, r.r,
// beta and p REDUNDANTLY in every CTA replaces the two scalar all-reduces + the published r of generations 1-3 by ONE exchange.
//   pull: owner threads store their rows' words into ONE global array [n]; every CTA polls all n entries (thread t: PL entries).
//   push: owners' values are stored into a private inbox of every CTA (G copies), each CTA polls only its own inbox.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;

template <int ST> __device__ __forceinline__ void put(u64 *p, u64 a, u64 b)
{
    if (ST == 0) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (ST == 1) { asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
                   asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p + 1), "l"(b) : "memory"); }
}
// two adjacent entries (32 bytes)
template <int LD> __device__ __forceinline__ void get2(const u64 *p, u64 (&w)[4])
{
    if (LD == 0) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
                   asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory"); }
    if (LD == 1) asm volatile("ld.relaxed.gpu.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p) : "memory");
    if (LD == 2) { asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
                   asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory"); }
    if (LD == 3) { asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
                   asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory"); }
}

// thread t of every CTA needs entries {w*64*K + 2*lane + 64*k, +1 : k < K} (K pairs) — the register layout of p in the solver
template <int ST, int LD, int K, bool PUSH>
__global__ void __launch_bounds__(512, 1) allgather(u64 *buf, int n, int rounds, long long *cycles, double *out)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, G = gridDim.x, bid = blockIdx.x;
    const int base = n / G, rem = n % G;
    const int r0 = bid * base + (bid < rem ? bid : rem), rcnt = base + (bid < rem ? 1 : 0);
    const int cbase = warp * 64 * K + 2 * lane;
    const size_t stride = (size_t)2 * n; // words per copy of the vector
    double acc = 0.0;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        const u64 tag = (u64)i << 32;
        u64 *cur = buf + (size_t)(i & 1) * (PUSH ? (size_t)G * stride : stride);
        // owners publish
        if (PUSH) {
            // 148 destinations x rcnt values: warp w serves destinations w, w + 16, ...; lane l < rcnt carries value l (contiguous 16-byte words)
            if (lane < rcnt) {
                const u64 bits = (u64)__double_as_longlong(1.0 + (r0 + lane) * 1e-3 + i);
                for (int d = warp; d < G; d += 16) put<ST>(cur + (size_t)d * stride + 2 * (size_t)(r0 + lane), tag | (bits >> 32), tag | (bits & 0xffffffffull));
            }
        } else if (t < rcnt) {
            const u64 bits = (u64)__double_as_longlong(1.0 + (r0 + t) * 1e-3 + i);
            put<ST>(cur + 2 * (size_t)(r0 + t), tag | (bits >> 32), tag | (bits & 0xffffffffull));
        }
        // everybody gathers
        const u64 *src = PUSH ? cur + (size_t)bid * stride : cur;
        u64 w[K][4];
        bool ok[K];
#pragma unroll
        for (int k = 0; k < K; ++k) ok[k] = cbase + 64 * k >= n; // nothing to fetch beyond n
        const long long tw = clock64();
        for (;;) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (!ok[k]) get2<LD>(src + 2 * (size_t)(cbase + 64 * k), w[k]);
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (!ok[k]) {
                    ok[k] = (w[k][0] >> 32) == (u64)i && (w[k][1] >> 32) == (u64)i && (w[k][2] >> 32) == (u64)i && (w[k][3] >> 32) == (u64)i;
                    all = all && ok[k];
                }
            if (all || clock64() - tw > 400000000LL) break;
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (cbase + 64 * k < n) {
                acc += __longlong_as_double((long long)(((w[k][0] & 0xffffffffull) << 32) | (w[k][1] & 0xffffffffull)));
                acc += __longlong_as_double((long long)(((w[k][2] & 0xffffffffull) << 32) | (w[k][3] & 0xffffffffull)));
            }
        __syncthreads();
    }
    if (t == 0 && bid == 0) *cycles = clock64() - t0;
    out[(size_t)bid * 512 + t] = acc;
}

template <int ST, int LD, int K, bool PUSH>
void run(const char *name, u64 *buf, size_t bytes, int n, int G, long long *d_cyc, double *d_out)
{
    int rounds = 1000;
    CK(cudaMemset(buf, 0, bytes));
    void *args[] = {&buf, &n, &rounds, &d_cyc, &d_out};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)allgather<ST, LD, K, PUSH>, dim3(G), dim3(512), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-58s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long cyc;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    static double host[148 * 512 + 512];
    CK(cudaMemcpy(host, d_out, (size_t)G * 512 * sizeof(double), cudaMemcpyDeviceToHost));
    double want = 0.0; // every CTA must have gathered the same totals: compare CTA 0 with the last CTA
    bool same = true;
    for (int t = 0; t < 512; ++t) same = same && host[t] == host[(size_t)(G - 1) * 512 + t];
    (void)want;
    printf("%-58s n=%5d  %6.0f cycles/round%s\n", name, n, (double)cyc / rounds, same ? "" : "  (CTAs DISAGREE)");
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    const size_t bytes = (size_t)2 * G * 2 * 4096 * sizeof(u64);
    u64 *buf;
    long long *d_cyc;
    double *d_out;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&d_cyc, sizeof(long long)));
    CK(cudaMalloc(&d_out, (size_t)G * 512 * sizeof(double)));
    printf("%s, %d SMs: all-gather of an n-vector as tagged words, cycles per round (incl. one CTA barrier)\n", prop.name, G);
    run<0, 0, 2, false>("pull  st.relaxed.gpu   / ld.relaxed.gpu.v2 x2", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 1, 2, false>("pull  st.relaxed.gpu   / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 2, 2, false>("pull  st.relaxed.gpu   / ld.cg.v2 x2 (weak)", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 3, 2, false>("pull  st.relaxed.gpu   / ld.volatile.v2 x2", buf, bytes, 2048, G, d_cyc, d_out);
    run<1, 0, 2, false>("pull  red.max x2       / ld.relaxed.gpu.v2 x2", buf, bytes, 2048, G, d_cyc, d_out);
    run<1, 1, 2, false>("pull  red.max x2       / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 0, 2, true>("push  st.relaxed.gpu   / ld.relaxed.gpu.v2 x2", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 1, 2, true>("push  st.relaxed.gpu   / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 2, 2, true>("push  st.relaxed.gpu   / ld.cg.v2 x2 (weak)", buf, bytes, 2048, G, d_cyc, d_out);
    run<0, 1, 1, false>("pull  st.relaxed.gpu   / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 1024, G, d_cyc, d_out);
    run<0, 1, 4, false>("pull  st.relaxed.gpu   / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 4096, G, d_cyc, d_out);
    run<0, 2, 4, false>("pull  st.relaxed.gpu   / ld.cg.v2 x2 (weak)", buf, bytes, 4096, G, d_cyc, d_out);
    run<0, 1, 4, true>("push  st.relaxed.gpu   / ld.relaxed.gpu.v4 (256-bit)", buf, bytes, 4096, G, d_cyc, d_out);
    return 0;
}
