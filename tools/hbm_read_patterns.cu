// hbm_read_patterns.cu — what read-only access pattern does B200's HBM3e reward?  (measurement tool, not product code)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hbm_read_patterns.out tools/hbm_read_patterns.cu
// Reads a 40 GB buffer with 16-byte streaming loads (ld.global.nc.L1::no_allocate.L2::evict_first) under different
// assignments of addresses to CTAs and reports GB/s (CUDA events, 3 passes after 1 warm-up; the buffer is >> L2):
//   chunk   : the buffer is cut into chunks of C bytes dealt to the CTAs round-robin (chunk id = blockIdx + k * grid), a CTA
//             reads its chunk front to back with U loads per thread in flight.  C = whole share  ->  one contiguous segment per
//             CTA; small C -> the whole GPU walks through memory almost sequentially (what a copy kernel does).
//   rows    : the GEMV's pattern — a CTA owns R rows of 800 KB at a time and sweeps them together, 32 KB per row per step.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ double2 ldg_stream(const double2 *p, unsigned long long pol)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int NT, int U>
__global__ void __launch_bounds__(NT) rd_chunks(const double2 *a, size_t n16, size_t chunk16, double *out)
{
    const unsigned long long pol = evict_first();
    double s = 0.0;
    for (size_t c0 = (size_t)blockIdx.x * chunk16; c0 < n16; c0 += (size_t)gridDim.x * chunk16) {
        const size_t c1 = c0 + chunk16 < n16 ? c0 + chunk16 : n16;
        size_t i = c0 + threadIdx.x;
        for (; i + (size_t)(U - 1) * NT < c1; i += (size_t)U * NT) {
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_stream(a + i + (size_t)u * NT, pol);
#pragma unroll
            for (int u = 0; u < U; ++u) s += v[u].x + v[u].y;
        }
        for (; i < c1; i += NT) {
            const double2 v = ldg_stream(a + i, pol);
            s += v.x + v.y;
        }
    }
    if (s == 12345.678) out[0] = s; // keep the loads alive
}

// R rows at a time, U loads of NT*16 bytes per row and step (the GEMV's sweep): row length row16 (16-byte units)
template <int NT, int R, int U>
__global__ void __launch_bounds__(NT) rd_rows(const double2 *a, size_t rows, size_t row16, double *out)
{
    const unsigned long long pol = evict_first();
    const size_t per = (rows + gridDim.x - 1) / gridDim.x;
    const size_t r_lo = (size_t)blockIdx.x * per, r_hi = r_lo + per < rows ? r_lo + per : rows;
    double s = 0.0;
    for (size_t r0 = r_lo; r0 + R <= r_hi; r0 += R) {
        for (size_t c = threadIdx.x; c + (size_t)(U - 1) * NT < row16; c += (size_t)U * NT) {
            double2 v[R][U];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u) v[r][u] = ldg_stream(a + (r0 + r) * row16 + c + (size_t)u * NT, pol);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u) s += v[r][u].x + v[r][u].y;
        }
    }
    if (s == 12345.678) out[0] = s;
}

static cudaEvent_t e0, e1;
template <typename F>
double time_gbs(F launch, double bytes)
{
    launch();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return bytes / (ms / 3 * 1e-3) / 1e9;
}

template <int NT, int U>
void sweep_chunks(const double2 *a, size_t n16, double *out, int sms)
{
    const double bytes = (double)n16 * 16;
    for (int per_sm : {1, 2, 4, 8}) {
        if (NT * per_sm > 2048) continue;
        const int grid = sms * per_sm;
        const size_t whole = (n16 + grid - 1) / grid;
        const size_t chunks[] = {(size_t)NT * U, (size_t)NT * U * 4, (size_t)NT * U * 16, whole};
        for (size_t c16 : chunks) {
            const double g = time_gbs([&] { rd_chunks<NT, U><<<grid, NT>>>(a, n16, c16, out); }, bytes);
            printf("chunk  NT=%4d U=%2d CTAs/SM=%d chunk=%10.1f KB : %7.1f GB/s\n", NT, U, per_sm, c16 * 16 / 1024.0, g);
        }
    }
}

template <int NT, int R, int U>
void sweep_rows(const double2 *a, size_t rows, size_t row16, double *out, int sms)
{
    for (int per_sm : {1, 2}) {
        if (NT * per_sm > 1024 && R * U > 16) continue;
        const int grid = sms * per_sm;
        const size_t per = (rows + grid - 1) / grid;
        const size_t used_rows = (per / R) * R * grid <= rows ? (per / R) * R * (size_t)grid : rows / R * R; // rows actually read
        const double g = time_gbs([&] { rd_rows<NT, R, U><<<grid, NT>>>(a, rows, row16, out); }, (double)used_rows * row16 * 16);
        printf("rows   NT=%4d R=%2d U=%2d CTAs/SM=%d                     : %7.1f GB/s\n", NT, R, U, per_sm, g);
    }
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t row16 = 49152, rows = 50000; // 98304 doubles per row = 786 KB (a multiple of every U*NT below); 50000 rows = 39.3 GB
    const size_t n16 = rows * row16;
    double2 *a;
    double *out;
    CK(cudaMalloc(&a, n16 * 16));
    CK(cudaMalloc(&out, 8));
    CK(cudaMemset(a, 0, n16 * 16));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("%s, %d SMs, %.1f GB buffer, read-only\n", prop.name, sms, n16 * 16 / 1e9);
    sweep_rows<512, 8, 4>(a, rows, row16, out, sms); // the default GEMV's shape
    sweep_rows<512, 4, 4>(a, rows, row16, out, sms);
    sweep_rows<512, 2, 8>(a, rows, row16, out, sms);
    sweep_rows<512, 1, 16>(a, rows, row16, out, sms);
    sweep_rows<256, 8, 4>(a, rows, row16, out, sms);
    sweep_rows<1024, 4, 4>(a, rows, row16, out, sms);
    sweep_chunks<256, 8>(a, n16, out, sms);
    sweep_chunks<256, 16>(a, n16, out, sms);
    sweep_chunks<512, 8>(a, n16, out, sms);
    sweep_chunks<512, 16>(a, n16, out, sms);
    sweep_chunks<512, 32>(a, n16, out, sms);
    sweep_chunks<1024, 8>(a, n16, out, sms);
    sweep_chunks<1024, 16>(a, n16, out, sms);
    return 0;
}
