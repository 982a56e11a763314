/*
 * ref_gpu_harness.cu — command-line window onto the UNMODIFIED reference GPU solver classes, so that the
 * reference's own CUDA path can be run on the same B200 as the product (BASELINE configs[1]: "n = 50000 on a
 * single B200 vs test_CG_single_GPU").
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/cg_oracle.c header): nothing under
 * 2024-eumaster4hpc-student-challenge_b200/ links, loads or executes this.  The file contains no reference code: it #includes the
 * reference's class header from where it lies under /root/reference and is linked by oracle/Makefile with the
 * reference's own .cu translation unit (compiled from where it lies, for sm_100) into
 *     oracle/_ref/ref_gpu_single.out   (-DREF_VARIANT=1: LAM::ConjugateGradient_GPU_CUDA<double>,
 *                                       ref: LAM/src/GPU/local/ConjugateGradient_GPU_CUDA.cu:225-316)
 *     oracle/_ref/ref_gpu_multi.out    (-DREF_VARIANT=2: LAM::ConjugateGradient_MultiGPUS_CUDA<double>, one process
 *                                       driving every visible GPU, ref: .../ConjugateGradient_MultiGPUS_CUDA.cu:225-470)
 * Separate executables (not one .so) because both reference units define identically named template kernels, and so that
 * the reference's out-of-bounds accesses (ref: GPU_CUDA.cu:22-31 warpReduce reads a[t+32..]) cannot touch the test process.
 *
 * Why a harness and not the reference's drivers (which oracle/Makefile also builds, unmodified): the drivers only take
 * files (a 20 GB file for n = 50000) and their solve() allocates, uploads A from pageable memory, iterates and frees in one
 * call, so the loop time can only be isolated by differencing runs with different max_iters on the SAME resident host
 * matrix.  `#define private public` lets this file point the class's private A/b/x/size at an in-memory system (the class
 * normally fills them in load_*_from_file, ref: GPU_CUDA.cu:318-380) and read x back, without touching the sources.
 *
 *   usage:  ref_gpu_X.out gen  <n> <rel_err> <x_out|-> K1 [K2 ...]          tridiag(1,2,1), b = 1 (generate mode)
 *           ref_gpu_X.out file <A.bin> <b.bin> <rel_err> <x_out|-> K1 [K2 ...]
 *   one JSON line per K on stdout: {"variant","n","max_iters","iters","rel","seconds","converged"}
 *   x of the LAST solve is written to x_out in the reference's file format (16-byte header + n doubles).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <iostream>
#include <memory>
#include <string>
#include <unistd.h>
#include <fcntl.h>
#include <cuda_runtime.h>

#define private public
#if REF_VARIANT == 1
#include "../src/GPU/local/ConjugateGradient_GPU_CUDA.cuh"
typedef LAM::ConjugateGradient_GPU_CUDA<double> RefSolver;
static const char *kVariant = "ConjugateGradient_GPU_CUDA<double>";
static const char *kTag = "PARALLEL GPU CUDA: ";
#else
#include "../src/GPU/local/ConjugateGradient_MultiGPUS_CUDA.cuh"
typedef LAM::ConjugateGradient_MultiGPUS_CUDA<double> RefSolver;
static const char *kVariant = "ConjugateGradient_MultiGPUS_CUDA<double>";
static const char *kTag = "PARALLEL MULTI-GPUS: ";
#endif
#undef private

namespace {

struct StdoutCapture {
    int saved_fd = -1;
    char path[64];
    StdoutCapture()
    {
        fflush(stdout);
        std::cout.flush();
        std::strcpy(path, "/tmp/lamcg_refgpu_XXXXXX");
        int fd = mkstemp(path);
        saved_fd = dup(1);
        dup2(fd, 1);
        close(fd);
    }
    std::string finish()
    {
        fflush(stdout);
        std::cout.flush();
        dup2(saved_fd, 1);
        close(saved_fd);
        std::string out;
        FILE *f = fopen(path, "rb");
        if (f) {
            char buf[4096];
            size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
            fclose(f);
        }
        unlink(path);
        return out;
    }
};

bool read_file(const char *path, size_t want_rows, size_t want_cols, double **out, size_t *rows)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
    size_t hdr[2];
    if (fread(hdr, sizeof(size_t), 2, f) != 2) { fclose(f); return false; }
    if ((want_cols && hdr[1] != want_cols) || (want_rows && hdr[0] != want_rows)) {
        fprintf(stderr, "%s: unexpected shape %zu x %zu\n", path, hdr[0], hdr[1]);
        fclose(f);
        return false;
    }
    size_t count = hdr[0] * hdr[1];
    double *buf = new double[count];
    size_t got = fread(buf, sizeof(double), count, f);
    fclose(f);
    if (got != count) { delete[] buf; return false; }
    *out = buf;
    *rows = hdr[0];
    return true;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s gen <n> <rel_err> <x_out|-> K1 [K2 ...]\n"
                        "       %s file <A.bin> <b.bin> <rel_err> <x_out|-> K1 [K2 ...]\n", argv[0], argv[0]);
        return 64;
    }
    int devs = 0;
    if (cudaGetDeviceCount(&devs) != cudaSuccess || devs == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    size_t n = 0;
    double *A = nullptr, *b = nullptr;
    int argi;
    if (std::strcmp(argv[1], "gen") == 0) {
        n = std::strtoull(argv[2], nullptr, 10);
        /* the reference's generate mode (ref: CPU/ConjugateGradient_CPU_MPI_OMP.hpp:144-256): 2 on the diagonal, 1 beside it, b = 1 */
        A = new double[n * n];
#pragma omp parallel for
        for (size_t i = 0; i < n; ++i) {
            double *row = A + i * n;
            std::memset(row, 0, n * sizeof(double));
            row[i] = 2.0;
            if (i > 0) row[i - 1] = 1.0;
            if (i + 1 < n) row[i + 1] = 1.0;
        }
        b = new double[n];
        for (size_t i = 0; i < n; ++i) b[i] = 1.0;
        argi = 3;
    } else if (std::strcmp(argv[1], "file") == 0) {
        if (argc < 7) return 64;
        size_t rows = 0;
        if (!read_file(argv[2], 0, 0, &A, &n)) return 1;
        if (!read_file(argv[3], n, 1, &b, &rows)) return 2;
        argi = 4;
    } else {
        return 64;
    }
    double rel_err = std::atof(argv[argi]);
    const char *x_path = argv[argi + 1];
    argi += 2;

    RefSolver cg;
    cg.A = A;
    cg.b = b;
    cg.x = new double[n];
    cg.size = n;

    cudaFree(0); /* context creation outside the timed call */
    for (; argi < argc; ++argi) {
        int K = std::atoi(argv[argi]);
        std::memset(cg.x, 0, n * sizeof(double));
        cudaDeviceSynchronize();
        StdoutCapture cap;
        auto t0 = std::chrono::high_resolution_clock::now();
        bool ok = cg.solve(K, rel_err);
        cudaDeviceSynchronize();
        auto t1 = std::chrono::high_resolution_clock::now();
        std::string out = cap.finish();
        cudaError_t err = cudaGetLastError();
        double seconds = std::chrono::duration<double>(t1 - t0).count();
        int iters = -1;
        double rel = NAN;
        size_t at = out.find(kTag);
        if (at != std::string::npos) {
            const char *s = out.c_str() + at + std::strlen(kTag);
            if (ok) sscanf(s, "Converged in %d iterations, relative error is %lf", &iters, &rel);
            else sscanf(s, "Did not converge in %d iterations, relative error is %lf", &iters, &rel);
        }
        printf("{\"variant\": \"%s\", \"devices\": %d, \"n\": %zu, \"max_iters\": %d, \"iters\": %d, \"rel\": %.17g, "
               "\"seconds\": %.6f, \"converged\": %s, \"cuda_error\": \"%s\"}\n",
               kVariant, REF_VARIANT == 1 ? 1 : devs, n, K, iters, std::isfinite(rel) ? rel : -1.0, seconds,
               ok ? "true" : "false", err == cudaSuccess ? "" : cudaGetErrorString(err));
        fflush(stdout);
    }
    if (std::strcmp(x_path, "-") != 0) {
        FILE *f = fopen(x_path, "wb");
        if (!f) return 6;
        size_t hdr[2] = {n, 1};
        fwrite(hdr, sizeof(size_t), 2, f);
        fwrite(cg.x, sizeof(double), n, f);
        fclose(f);
    }
    return 0;
}
