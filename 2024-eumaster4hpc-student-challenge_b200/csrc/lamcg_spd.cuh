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

// ---------------------------------------------------------------------------------------------
// The same product on the fp64 TENSOR cores (DMMA: mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) — the one dense contraction of this
// code base, hence the one place where tensor cores are legitimate.  128 x 128 output tile per 256-thread CTA,
// 8 warps as 2 (M) x 4 (N), each warp a 64 x 32 sub-tile = 8 x 4 fragments of 8 x 8 (64 accumulator doubles per thread), K step
// 16, operands double-buffered in shared memory as [k][m] / [k][n] with a row pitch of 132 doubles: 132 = 4 (mod 16) makes the
// 8-byte fragment loads (lane l reads [k0 + l % 4][m0 + l / 4]) hit 16 distinct 8-byte bank pairs per half warp.  The next K
// step's operands are fetched from global memory into registers while the current one is multiplied.  Arbitrary strides as in
// gemm_f64_kernel; the same deterministic split-K protocol (slice z writes its partial tile to ws, summed in slice order).
// ---------------------------------------------------------------------------------------------
constexpr int kMmaTM = 128, kMmaTN = 128, kMmaTK = 16, kMmaPitch = 132, kMmaThreads = 256;
constexpr size_t kMmaSmemBytes = (size_t)2 * 2 * kMmaTK * kMmaPitch * sizeof(double); // A and B, two stages each

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kMmaThreads, 1) gemm_f64_mma_kernel(GemmArgs g)
{
    constexpr int TM = kMmaTM, TN = kMmaTN, TK = kMmaTK, P = kMmaPitch;
    extern __shared__ __align__(16) double gsm[];
    double *As = gsm;               // [2][TK][P]
    double *Bs = gsm + 2 * TK * P;  // [2][TK][P]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32; // this warp's sub-tile origin inside the CTA tile
    const int fr = lane >> 2, fc = lane & 3;               // fragment coordinates: row / column index l/4, k index l%4
    const long long i0 = (long long)blockIdx.y * TM, j0 = (long long)blockIdx.x * TN;
    const long long kb = g.ws ? (long long)blockIdx.z * g.k_slice : 0;
    const long long ke = g.ws ? (kb + g.k_slice < g.K ? kb + g.k_slice : g.K) : g.K;

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // staging: TM x TK (and TK x TN) = 2048 elements per operand, 8 per thread; the index split walks the unit-stride
    // dimension of the operand fastest across threads (coalesced global reads for every N/T combination)
    double ra[8], rb[8];
    auto fetch = [&](long long k0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * kMmaThreads;
            int am, ak;
            if (g.sai == 1) { am = idx % TM; ak = idx / TM; } else { ak = idx % TK; am = idx / TK; }
            const long long gi = i0 + am, gk = k0 + ak;
            ra[e] = (gi < g.M && gk < ke) ? g.A[gi * g.sai + gk * g.sak] : 0.0;
            int bn, bk;
            if (g.sbj == 1) { bn = idx % TN; bk = idx / TN; } else { bk = idx % TK; bn = idx / TK; }
            const long long gj = j0 + bn, gk2 = k0 + bk;
            rb[e] = (gj < g.N && gk2 < ke) ? g.B[gk2 * g.sbk + gj * g.sbj] : 0.0;
        }
    };
    auto stash = [&](int stage) {
        double *as = As + stage * TK * P, *bs = Bs + stage * TK * P;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * kMmaThreads;
            int am, ak;
            if (g.sai == 1) { am = idx % TM; ak = idx / TM; } else { ak = idx % TK; am = idx / TK; }
            as[ak * P + am] = ra[e];
            int bn, bk;
            if (g.sbj == 1) { bn = idx % TN; bk = idx / TN; } else { bk = idx % TK; bn = idx / TK; }
            bs[bk * P + bn] = rb[e];
        }
    };

    if (kb < ke) {
        fetch(kb);
        stash(0);
    }
    __syncthreads();
    int stage = 0;
    for (long long k0 = kb; k0 < ke; k0 += TK, stage ^= 1) {
        const bool more = k0 + TK < ke;
        if (more) fetch(k0 + TK); // global loads in flight while this stage is multiplied
        const double *as = As + stage * TK * P, *bs = Bs + stage * TK * P;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int a = 0; a < 8; ++a) af[a] = as[(kk + fc) * P + wm + 8 * a + fr];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = bs[(kk + fc) * P + wn + 8 * b + fr];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        if (more) stash(stage ^ 1); // the other stage was last read before the barrier that ended the previous step
        __syncthreads();
    }
    // C fragment of m8n8k4: lane l holds row l/4, columns 2*(l%4) and 2*(l%4)+1
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long gi = i0 + wm + 8 * a + fr;
        if (gi >= g.M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const long long gj = j0 + wn + 8 * b + 2 * fc + q;
                if (gj >= g.N) continue;
                const double v = acc[a][b][q];
                if (g.ws) {
                    g.ws[(long long)blockIdx.z * g.M * g.N + gi * g.N + gj] = v;
                    continue;
                }
                double *c = g.C + gi * g.sci + gj * g.scj;
                *c = g.beta == 0.0 ? g.alpha * v : fma(g.alpha, v, g.beta * *c);
            }
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

// ---------------------------------------------------------------------------------------------
// glibc's rand() stream on the device.  random_spd_system.cpp:27-38 fills the n x n matrix with 2*rand()/RAND_MAX - 1 after
// srand(seed); calling rand() n^2 times on the host costs ~25 ns each (a lock per call: 6-7 s at n = 16384, most of the
// generator's time in round 1).  glibc's default generator (TYPE_3, stdlib/random_r.c) is the additive feedback recurrence
//     v[k] = v[k-3] + v[k-31]  (mod 2^32),      rand() = v[k] >> 1,
// which is LINEAR: the host jumps ahead with powers of its 31 x 31 companion matrix and hands every thread the 31-word state at
// the start of its chunk (lamcg.cu: glibc_states); the thread then produces its chunk sequentially.  Bit-identical to the host
// stream (tests/test_gpu_spd_generator.py compares with a rand()-based fill on the host).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) glibc_rand_fill_kernel(double *out, long long count, const unsigned int *states, long long chunk)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long first = t * chunk;
    if (first >= count) return;
    unsigned int v[31]; // v[j] = element k + j of the sequence, k = position of this thread's next output minus 31
#pragma unroll
    for (int j = 0; j < 31; ++j) v[j] = states[t * 31 + j];
    const long long last = first + chunk < count ? first + chunk : count;
    for (long long i0 = first; i0 < last; i0 += 31) {
#pragma unroll
        for (int j = 0; j < 31; ++j) { // static ring indices: v[j] becomes element k + 31 + j = (k + j) + (k + 28 + j)
            v[j] += v[(j + 28) % 31];
            if (i0 + j < last) out[i0 + j] = ((2.0 * (double)(int)(v[j] >> 1)) / 2147483647.0) - 1.0;
        }
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

// Leaf PANEL of the Gram-Schmidt recursion (w <= 32 columns).  The reference recurses down to single columns
// (random_spd_system.cpp:41-62); here the last five levels (up to 31 projections + 32 normalisations = ~125 tiny launches per
// panel, 65 k launches at n = 16384: the generator was launch bound) are replaced by CholeskyQR2 on the n x w panel: G = P^T P,
// G = R^T R, P <- P R^-1, twice.  The thin QR factor with a positive diagonal of R is unique, so this is the same Q that
// Gram-Schmidt produces, up to rounding (the second pass brings the orthogonality error to the level of the working precision).
// This kernel is the middle step: one warp, lane j owns column j of R and of R^-1 (w x w, column-major, ld = w, in and out).
__global__ void __launch_bounds__(32) chol_inverse_kernel(const double *G, int w, double *Rinv)
{
    __shared__ double R[32][33]; // R[i][j], upper triangular
    __shared__ double X[32][33]; // R^-1
    const int j = threadIdx.x;
    for (int i = 0; i < 32; ++i) { R[i][j] = 0.0; X[i][j] = 0.0; }
    __syncwarp();
    // Cholesky, row by row: R[k][k] = sqrt(G[k][k] - sum_{i<k} R[i][k]^2); R[k][j] = (G[k][j] - sum_{i<k} R[i][k] R[i][j]) / R[k][k]
    for (int k = 0; k < w; ++k) {
        if (j >= k && j < w) {
            double acc = G[(size_t)j * w + k]; // G is symmetric
            for (int i = 0; i < k; ++i) acc = fma(-R[i][k], R[i][j], acc);
            R[k][j] = acc; // not yet divided
        }
        __syncwarp();
        const double d = sqrt(R[k][k]);
        __syncwarp();
        if (j >= k && j < w) R[k][j] = (j == k) ? d : R[k][j] / d;
        __syncwarp();
    }
    // X = R^-1 by back substitution, lane j solves R x = e_j
    if (j < w) {
        X[j][j] = 1.0 / R[j][j];
        for (int i = j - 1; i >= 0; --i) {
            double acc = 0.0;
            for (int k = i + 1; k <= j; ++k) acc = fma(R[i][k], X[k][j], acc);
            X[i][j] = -acc / R[i][i];
        }
    }
    __syncwarp();
    if (j < w)
        for (int i = 0; i < w; ++i) Rinv[(size_t)j * w + i] = X[i][j];
}

// Column c of the column-major n x n matrix *= sqrt(d[c])   (cblas_dscal per column, :89-92)
__global__ void __launch_bounds__(256) scale_columns_kernel(double *Q, const double *d, long long n)
{
    const long long total = n * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        Q[t] *= sqrt(d[t / n]);
}

} // namespace lamcgk
