/*
 * ref_harness.cpp — C-ABI window onto the UNMODIFIED reference solver classes.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/cg_oracle.c header).  This file contains no reference
 * code: it #includes the reference's LAM.hpp from where it lies under /root/reference (the path
 * is given on the compiler command line by oracle/Makefile) and is compiled into
 * oracle/_ref/libref_harness.so, which is git-ignored and travels to the GPU box prebuilt.
 *
 * Why a harness: ConjugateGradient_CPU_MPI_OMP::save_result_to_file writes the rhs instead of
 * the solution (MPI_OMP.hpp:436-439), and neither CPU class can be fed an in-memory system, so
 * the only way to observe the reference's x and to time its solve() without file I/O is to
 * reach its private members.  `#define private public` does that without touching the sources.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <iostream>
#include <string>
#include <unistd.h>
#include <fcntl.h>
#include <omp.h>

#define private public
#include "LAM.hpp"
#undef private

namespace {

/* The reference prints its results on stdout from inside solve(); capture them. */
struct StdoutCapture {
    int saved_fd = -1;
    char path[64];
    StdoutCapture()
    {
        fflush(stdout);
        std::cout.flush();
        std::strcpy(path, "/tmp/lamcg_ref_XXXXXX");
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

} // namespace

extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/*
 * Generate mode through LAM::ConjugateGradient_CPU_MPI_OMP<double> (1 rank via the mpi shim):
 * generate_matrix(n,n), generate_rhs(), solve(max_iters, rel_err).  Outputs: x (n doubles,
 * nullable), iteration count exactly as the class prints it (max_iters+1 when not converged),
 * relative residual parsed from the class's own CSV fields, and wall seconds of solve() alone and
 * of generate_matrix() alone.  Returns solve()'s bool (1/0), -1 on parse failure.
 */
int ref_gen_solve(size_t n, int max_iters, double rel_err, double *x_out, int *iters_out,
                  double *rel_out, double *solve_seconds, double *gen_seconds)
{
    LAM::ConjugateGradient_CPU_MPI_OMP<double> cg;
    StdoutCapture cap;
    double t0 = omp_get_wtime();
    cg.generate_matrix(n, n);
    double t1 = omp_get_wtime();
    cg.generate_rhs();
    double t2 = omp_get_wtime();
    bool ok = cg.solve(max_iters, rel_err);
    double t3 = omp_get_wtime();
    std::string out = cap.finish();
    if (gen_seconds) *gen_seconds = t1 - t0;
    if (solve_seconds) *solve_seconds = t3 - t2;
    /* stdout so far: "<n>,<avg_gemv>,<avg_iter>,<iters>,<rel>," */
    unsigned long nn = 0;
    double g = 0, it = 0, rel = 0;
    int iters = 0;
    int got = sscanf(out.c_str(), "%lu,%lf,%lf,%d,%lf,", &nn, &g, &it, &iters, &rel);
    if (x_out) std::memcpy(x_out, cg._x, n * sizeof(double));
    delete[] cg._matrix; delete[] cg._rhs; delete[] cg._x; delete[] cg._r; delete[] cg._Ap; delete[] cg._p;
    delete[] cg._sendcounts; delete[] cg._displs;
    if (got != 5) return -1;
    if (iters_out) *iters_out = iters;
    if (rel_out) *rel_out = rel;
    return ok ? 1 : 0;
}

/*
 * The same generate-mode path with the object kept alive between solves (bench.py --impl reference at full size: the 80 GB
 * matrix of n = 100000 is generated once, every timed step is one solve() of a few iterations on it; the reference's solve()
 * re-initialises x, r, p itself, MPI_OMP.hpp:82-89).  ref_gen_open returns an opaque handle (NULL when the allocation fails).
 */
void *ref_gen_open(size_t n, double *gen_seconds)
{
    LAM::ConjugateGradient_CPU_MPI_OMP<double> *cg = nullptr;
    StdoutCapture cap;
    double t0 = omp_get_wtime();
    bool ok = false;
    try {
        cg = new LAM::ConjugateGradient_CPU_MPI_OMP<double>();
        ok = cg->generate_matrix(n, n) && cg->generate_rhs();
    } catch (...) { /* std::bad_alloc from the reference's new[] */
        ok = false;
    }
    double t1 = omp_get_wtime();
    cap.finish();
    if (gen_seconds) *gen_seconds = t1 - t0;
    if (!ok) {
        delete cg;
        return nullptr;
    }
    return cg;
}

int ref_gen_solve_again(void *h, int max_iters, double rel_err, double *x_out, int *iters_out, double *rel_out, double *solve_seconds)
{
    auto *cg = static_cast<LAM::ConjugateGradient_CPU_MPI_OMP<double> *>(h);
    if (!cg) return -1;
    StdoutCapture cap;
    double t0 = omp_get_wtime();
    bool ok = cg->solve(max_iters, rel_err);
    double t1 = omp_get_wtime();
    std::string out = cap.finish();
    if (solve_seconds) *solve_seconds = t1 - t0;
    double g = 0, it = 0, rel = 0;
    int iters = 0;
    int got = sscanf(out.c_str(), "%lf,%lf,%d,%lf,", &g, &it, &iters, &rel); /* "<avg_gemv>,<avg_iter>,<iters>,<rel>," */
    if (x_out) std::memcpy(x_out, cg->_x, cg->_num_cols * sizeof(double));
    if (got != 4) return -1;
    if (iters_out) *iters_out = iters;
    if (rel_out) *rel_out = rel;
    return ok ? 1 : 0;
}

void ref_gen_close(void *h)
{
    auto *cg = static_cast<LAM::ConjugateGradient_CPU_MPI_OMP<double> *>(h);
    if (!cg) return;
    delete[] cg->_matrix; delete[] cg->_rhs; delete[] cg->_x; delete[] cg->_r; delete[] cg->_Ap; delete[] cg->_p;
    delete[] cg->_sendcounts; delete[] cg->_displs;
    delete cg;
}

/*
 * In-memory system through LAM::ConjugateGradient_CPU_OMP<double>::solve.  The class normally
 * fills its members in load_matrix_from_file/load_rhs_from_file (OMP.hpp:93-197); here they are
 * pointed at the caller's buffers instead, which leaves solve() itself untouched.
 */
int ref_omp_solve(const double *A, const double *b, size_t n, int max_iters, double rel_err,
                  double *x_out, int *iters_out, double *rel_out, double *solve_seconds)
{
    LAM::ConjugateGradient_CPU_OMP<double> cg;
    cg._num_rows = n;
    cg._num_cols = n;
    cg._matrix = const_cast<double *>(A);
    cg._rhs = const_cast<double *>(b);
    cg._x = x_out;
    cg._r = new double[n];
    cg._p = new double[n];
    cg._Ap = new double[n];
    StdoutCapture cap;
    double t0 = omp_get_wtime();
    bool ok = cg.solve(max_iters, rel_err);
    double t1 = omp_get_wtime();
    std::string out = cap.finish();
    if (solve_seconds) *solve_seconds = t1 - t0;
    delete[] cg._r; delete[] cg._p; delete[] cg._Ap;
    int iters = 0;
    double rel = 0;
    int got;
    if (ok) got = sscanf(out.c_str(), "Converged in %d iterations, relative error is %lf", &iters, &rel);
    else got = sscanf(out.c_str(), "Did not converge in %d iterations, relative error is %lf", &iters, &rel);
    if (got != 2) return -1;
    if (iters_out) *iters_out = iters;
    if (rel_out) *rel_out = rel;
    return ok ? 1 : 0;
}

/* The <float> instantiation of the same reference class (in-memory system), for the fp32 path. */
int ref_omp_solve_f32(const float *A, const float *b, size_t n, int max_iters, float rel_err, float *x_out, int *iters_out,
                      float *rel_out)
{
    LAM::ConjugateGradient_CPU_OMP<float> cg;
    cg._num_rows = n;
    cg._num_cols = n;
    cg._matrix = const_cast<float *>(A);
    cg._rhs = const_cast<float *>(b);
    cg._x = x_out;
    cg._r = new float[n];
    cg._p = new float[n];
    cg._Ap = new float[n];
    StdoutCapture cap;
    bool ok = cg.solve(max_iters, rel_err);
    std::string out = cap.finish();
    delete[] cg._r; delete[] cg._p; delete[] cg._Ap;
    int iters = 0;
    float rel = 0;
    int got;
    if (ok) got = sscanf(out.c_str(), "Converged in %d iterations, relative error is %f", &iters, &rel);
    else got = sscanf(out.c_str(), "Did not converge in %d iterations, relative error is %f", &iters, &rel);
    if (got != 2) return -1;
    if (iters_out) *iters_out = iters;
    if (rel_out) *rel_out = rel;
    return ok ? 1 : 0;
}

} // extern "C"
