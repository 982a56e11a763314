// mixed_probe.cu — microbenchmark (measurement tool, not product code): a row-sweep GEMV whose matrix is STORED in fp32 while p,
// the products and the row sums stay fp64 (SURVEY §8(f)-3 "fp32-A / fp64-accumulate").
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mixed_probe.out tools/mixed_probe.cu && tools/mixed_probe.out
// Question: the fp32 handle's first version converted every fp32 product to fp64 and ran no faster than the fp64 sweep at half the
// bytes.  Is the fp32 -> fp64 conversion instruction (F2F.F64.F32) the limiter, and does building the double from the float's
// bits with integer instructions (exact for normal numbers and zero) lift it back onto the HBM roofline?
//   CONV 0: cvt.f64.f32 per element     CONV 1: integer bit construction per element
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ double widen_cvt(float f) { return (double)f; }
// normal numbers and +-0 only: sign | (exponent + 896) | mantissa << 29
__device__ __forceinline__ double widen_bits(float f)
{
    const unsigned b = __float_as_uint(f);
    const unsigned mag = b & 0x7fffffffu;
    unsigned hi = (mag >> 3) + (mag ? 0x38000000u : 0u);
    hi |= b & 0x80000000u;
    return __hiloint2double((int)hi, (int)(b << 29));
}
template <int CONV> __device__ __forceinline__ double widen(float f) { return CONV ? widen_bits(f) : widen_cvt(f); }

__device__ __forceinline__ float4 ldg_stream(const float *p, uint64_t pol)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// R rows per pass, thread t owns 4 consecutive columns out of every NT*4 of a NT*U*4-column chunk (cols % chunk == 0 here)
template <int R, int U, int NT, int CPS, int CONV>
__global__ void __launch_bounds__(NT, CPS) gemv_mixed(const float *__restrict__ A, const double *__restrict__ p, double *__restrict__ y,
                                                      long long rows, long long n)
{
    __shared__ double red[NT / 32][R];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long per = rows / gridDim.x; // multiple of R
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    constexpr int CH = NT * U * 4;
    for (long long rs = blockIdx.x * per; rs < (blockIdx.x + 1) * per; rs += R) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        const float *pa = A + rs * n + 4 * tid;
        const double *pp = p + 4 * tid;
        for (long long c = 0; c < n; c += CH, pa += CH, pp += CH) {
            double pv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2 lo = __ldg(reinterpret_cast<const double2 *>(pp + u * NT * 4));
                const double2 hi = __ldg(reinterpret_cast<const double2 *>(pp + u * NT * 4 + 2));
                pv[u][0] = lo.x; pv[u][1] = lo.y; pv[u][2] = hi.x; pv[u][3] = hi.y;
            }
            float4 a[R][U];
            const float *pr = pa;
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int u = 0; u < U; ++u) a[r][u] = ldg_stream(pr + u * NT * 4, pol);
                pr += n;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[r] = __dadd_rn(__dmul_rn(widen<CONV>(a[r][u].x), pv[u][0]), acc[r]);
                    acc[r] = __dadd_rn(__dmul_rn(widen<CONV>(a[r][u].y), pv[u][1]), acc[r]);
                    acc[r] = __dadd_rn(__dmul_rn(widen<CONV>(a[r][u].z), pv[u][2]), acc[r]);
                    acc[r] = __dadd_rn(__dmul_rn(widen<CONV>(a[r][u].w), pv[u][3]), acc[r]);
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
        if (warp == 0 && lane < R) {
            double s = 0.0;
            for (int w = 0; w < NT / 32; ++w) s += red[w][lane];
            y[rs + lane] = s;
        }
        __syncthreads();
    }
}

__global__ void fill(float *A, long long count, double *p, long long n)
{
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < count; i += step) {
        unsigned h = (unsigned)(i * 2654435761u) ^ (unsigned)(i >> 32);
        h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
        A[i] = (h & 7u) == 0 ? 0.0f : (float)((int)(h >> 8) - (1 << 23)) * (1.0f / (1 << 20));
    }
    for (long long i = i0; i < n; i += step) p[i] = 1.0 + 1e-3 * (double)(i % 977);
}

// widen_bits against the conversion instruction over every normal / zero bit pattern with the given low bits step
__global__ void check_widen(unsigned long long *bad)
{
    const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    for (unsigned long long b = i0; b < (1ull << 32); b += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned e = ((unsigned)b >> 23) & 0xffu;
        if ((e == 0 && ((unsigned)b & 0x7fffffu)) || e == 255) continue; // denormal, inf, nan: not covered by construction
        const float f = __uint_as_float((unsigned)b);
        if (__double_as_longlong(widen_bits(f)) != __double_as_longlong((double)f)) ++mine;
    }
    if (mine) atomicAdd(bad, mine);
}

template <int R, int U, int NT, int CPS, int CONV>
static void run(const char *name, const float *A, const double *p, double *y, long long rows, long long n, int sms, double *ref)
{
    const int grid = sms * CPS;
    const long long use = rows / ((long long)grid * R) * grid * R;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) gemv_mixed<R, U, NT, CPS, CONV><<<grid, NT>>>(A, p, y, use, n);
    CK(cudaDeviceSynchronize());
    const int reps = 5;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) gemv_mixed<R, U, NT, CPS, CONV><<<grid, NT>>>(A, p, y, use, n);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    double *h = (double *)malloc(use * sizeof(double));
    CK(cudaMemcpy(h, y, use * sizeof(double), cudaMemcpyDeviceToHost));
    long long diff = 0;
    if (ref[0] == -1.0) { for (long long i = 0; i < use; ++i) ref[i + 1] = h[i]; ref[0] = (double)use; }
    else for (long long i = 0; i < use && i < (long long)ref[0]; ++i) diff += h[i] != ref[i + 1];
    printf("%-20s rows %lld  %.3f ms  %.0f GB/s of fp32 matrix  y[1]=%.17g  rows_differing_from_first_run=%lld\n", name, use, ms,
           4.0 * use * n / ms * 1e-6, h[1], diff);
    free(h);
}

int main(int argc, char **argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 98304; // multiple of every chunk size used below (512*4*4 = 8192)
    const long long rows = argc > 2 ? atoll(argv[2]) : 47360;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    float *A; double *p, *y;
    CK(cudaMalloc(&A, rows * n * sizeof(float)));
    CK(cudaMalloc(&p, n * sizeof(double)));
    CK(cudaMalloc(&y, rows * sizeof(double)));
    fill<<<prop.multiProcessorCount * 8, 256>>>(A, rows * n, p, n);
    unsigned long long *bad;
    CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    check_widen<<<prop.multiProcessorCount * 8, 256>>>(bad);
    unsigned long long hbad = 0;
    CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
    printf("widen_bits vs cvt.f64.f32 over all normal and zero fp32 bit patterns: %llu mismatches\n", hbad);
    printf("n %lld, %.1f GB of fp32 matrix, %d SMs\n", n, rows * n * 4e-9, prop.multiProcessorCount);
    double *ref = (double *)malloc((rows + 1) * sizeof(double));
    ref[0] = -1.0;
    run<8, 4, 512, 1, 0>("cvt   R8 U4 512x1", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 4, 512, 1, 1>("bits  R8 U4 512x1", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 2, 512, 1, 0>("cvt   R8 U2 512x1", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 2, 512, 1, 1>("bits  R8 U2 512x1", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 4, 256, 2, 0>("cvt   R8 U4 256x2", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 4, 256, 2, 1>("bits  R8 U4 256x2", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<8, 2, 256, 2, 1>("bits  R8 U2 256x2", A, p, y, rows, n, prop.multiProcessorCount, ref);
    run<4, 4, 512, 1, 1>("bits  R4 U4 512x1", A, p, y, rows, n, prop.multiProcessorCount, ref);
    return 0;
}
