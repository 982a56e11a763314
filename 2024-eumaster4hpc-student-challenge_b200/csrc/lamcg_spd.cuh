// lamcg_spd.cuh — device side of the random SPD system generator.
//
// Reference: challenge/main/random_spd_system.cpp (Intel MKL on the host): Q = recursive block
// Gram-Schmidt of a U(-1,1) matrix (:41-62), D = exp(3.5 U(-1,1)) (:83-87), A = (Q sqrt(D))(Q sqrt(D))^T
// (:89-96), rhs U(-1,1) (:166).  Here the random streams stay on the host (they are glibc rand(),
// which is what makes a seed reproduce the reference's draws) and every O(n^3) step runs on the GPU:
// one strided fp64 GEMM kernel serves Q1^T Q2, Q2 -= Q1 C and Y Y^T.  This is a setup tool, not the
// hot path; fused multiply-add is used (the generator's output is "parity unpinned": no reference
// test pins MKL's bits).
#pragma once

#include "lamcg_device.cuh"

namespace lamcgk {

// C[i*sci + j*scj] = alpha * sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj] + beta * C[...]
// 64x64 output tile per 256-thread CTA, 4x4 outputs per thread, K step 16, operands staged in
// shared memory.  Arbitrary strides cover the N/T combinations of row- and column-major storage.
struct GemmArgs {
    const double *A, *B;
    double *C;
    long long M, N, K;
    long long sai, sak, sbk, sbj, sci, scj;
    double alpha, beta;
    long long k_slice;  // split-K: CTA z handles K range [z*k_slice, min(K, (z+1)*k_slice)) and writes its partial
    double *ws;         //          product (alpha = 1, beta = 0, compact row-major M x N) to ws + z*M*N
};

__global__ void __launch_bounds__(256) gemm_f64_kernel(GemmArgs g)
{
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4; // 16 x 16 threads, each 4 x 4 outputs
    const long long i0 = (long long)blockIdx.y * TM, j0 = (long long)blockIdx.x * TN;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const long long kb = g.ws ? (long long)blockIdx.z * g.k_slice : 0;
    const long long ke = g.ws ? (kb + g.k_slice < g.K ? kb + g.k_slice : g.K) : g.K;
    for (long long k0 = kb; k0 < ke; k0 += TK) {
        // stage A tile (TM x TK) and B tile (TK x TN): 1024 elements each, 4 per thread
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            // choose the index split that walks the unit-stride dimension fastest across threads
            int am, ak;
            if (g.sai == 1) { am = idx % TM; ak = idx / TM; } else { ak = idx % TK; am = idx / TK; }
            const long long gi = i0 + am, gk = k0 + ak;
            As[ak][am] = (gi < g.M && gk < ke) ? g.A[gi * g.sai + gk * g.sak] : 0.0;
            int bn, bk;
            if (g.sbj == 1) { bn = idx % TN; bk = idx / TN; } else { bk = idx % TK; bn = idx / TK; }
            const long long gj = j0 + bn, gk2 = k0 + bk;
            Bs[bk][bn] = (gj < g.N && gk2 < ke) ? g.B[gk2 * g.sbk + gj * g.sbj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[k][ty * 4 + q];
                b[q] = Bs[k][tx * 4 + q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const long long gi = i0 + ty * 4 + p;
        if (gi >= g.M) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long gj = j0 + tx * 4 + q;
            if (gj >= g.N) continue;
            if (g.ws) {
                g.ws[(long long)blockIdx.z * g.M * g.N + gi * g.N + gj] = acc[p][q];
                continue;
            }
            double *c = g.C + gi * g.sci + gj * g.scj;
            *c = g.beta == 0.0 ? g.alpha * acc[p][q] : fma(g.alpha, acc[p][q], g.beta * *c);
        }
    }
}

// Second pass of a split-K product: C = alpha * (sum of the slices, in slice order) + beta * C.
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(GemmArgs g, int slices)
{
    const long long total = g.M * g.N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int z = 0; z < slices; ++z) s += g.ws[(long long)z * total + t];
        const long long i = t / g.N, j = t - i * g.N;
        double *c = g.C + i * g.sci + j * g.scj;
        *c = g.beta == 0.0 ? g.alpha * s : fma(g.alpha, s, g.beta * *c);
    }
}

// Leaf of the Gram-Schmidt recursion: x /= ||x||_2 for one column (cblas_dnrm2 + cblas_dscal).
__global__ void __launch_bounds__(256) normalize_column_kernel(double *x, long long n)
{
    __shared__ double scratch[32];
    __shared__ double s_inv;
    double local = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) local = fma(x[i], x[i], local);
    const double ss = block_sum(local, scratch);
    if (threadIdx.x == 0) s_inv = 1.0 / sqrt(ss);
    __syncthreads();
    const double inv = s_inv;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) x[i] *= inv;
}

// Column c of the column-major n x n matrix *= sqrt(d[c])   (cblas_dscal per column, :89-92)
__global__ void __launch_bounds__(256) scale_columns_kernel(double *Q, const double *d, long long n)
{
    const long long total = n * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        Q[t] *= sqrt(d[t / n]);
}

} // namespace lamcgk
