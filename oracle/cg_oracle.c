/*
 * cg_oracle.c — CPU restatement of the reference's dense CG path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker.  The product (lamcg CUDA library) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (a) the golden vectors
 * in tests/golden/ that were produced by the unmodified reference sources compiled from
 * /root/reference (oracle/Makefile -> oracle/_ref/, generator tests/golden/make_golden.py), and
 * (b) the known answers in the reference's own result dumps (TESTS/BEST_RESULTS:173,184,214,
 * TESTS/results/STRESS_TEST_GPU_MPI.txt:17-18, TESTS/results/WEAK_SCALABILITY_GPU_MPI.txt:20).
 *
 * All "ref:" citations are into /root/reference/challenge/main/LAM/src/.
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA), so
 * a*b+c is two roundings there.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ref: CPU/ConjugateGradient_CPU_OMP.hpp:219-231 (dot).  The reference combines OpenMP thread
 * partials in an unspecified order; the oracle fixes the order to plain left-to-right, which is
 * what the reference computes with OMP_NUM_THREADS=1. */
double oracle_dot(const double *x, const double *y, size_t size)
{
    double result = 0.0;
    for (size_t i = 0; i < size; i++) result += x[i] * y[i];
    return result;
}

/* ref: CPU/ConjugateGradient_CPU_OMP.hpp:233-244 (axpby): y = alpha*x + beta*y, two products
 * and one sum per element, no contraction. */
void oracle_axpby(double alpha, const double *x, double beta, double *y, size_t size)
{
    for (size_t i = 0; i < size; i++) y[i] = alpha * x[i] + beta * y[i];
}

/* ref: CPU/ConjugateGradient_CPU_OMP.hpp:246-263 (gemv); distributed form
 * CPU/ConjugateGradient_CPU_MPI_OMP.hpp:482-508.  Per row: strictly sequential left-to-right
 * accumulation of (alpha*A[r,c])*x[c], then y[r] = beta*y[r] + y_val.  Rows are independent, so
 * the row loop may run in parallel without changing a single bit. */
void oracle_gemv(double alpha, const double *A, const double *x, double beta, double *y,
                 size_t rows, size_t cols)
{
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; r++) {
        double y_val = 0.0;
        const double *a = A + r * cols;
        for (size_t c = 0; c < cols; c++) y_val += alpha * a[c] * x[c];
        y[r] = beta * y[r] + y_val;
    }
}

/* ref: CPU/ConjugateGradient_CPU_MPI_OMP.hpp:175-196 (row partition): n/P rows each, the
 * remainder goes to the last rank; offset = rank*(n/P). */
void oracle_partition(size_t n, int nranks, int rank, size_t *rows, size_t *offset)
{
    size_t base = n / (size_t)nranks;
    *offset = base * (size_t)rank;
    *rows = base + ((rank == nranks - 1) ? n % (size_t)nranks : 0);
}

/* ref: CPU/ConjugateGradient_CPU_MPI_OMP.hpp:237-247 (generate-mode matrix): for the local row i
 * with global row g = i + offset: 1 on the two off-diagonals, 2 on the diagonal, 0 elsewhere. */
void oracle_generate_matrix(double *A, size_t local_rows, size_t cols, size_t offset)
{
    for (size_t i = 0; i < local_rows; i++) {
        for (size_t j = 0; j < cols; j++) {
            size_t g = i + offset;
            double v;
            if (g == j - 1 || g == j + 1) v = 1.0;
            else if (g == j) v = 2.0;
            else v = 0.0;
            A[i * cols + j] = v;
        }
    }
}

/* ref: CPU/ConjugateGradient_CPU_MPI_OMP.hpp:159-162 (generate-mode rhs): all ones. */
void oracle_generate_rhs(double *b, size_t n)
{
    for (size_t i = 0; i < n; i++) b[i] = 1.0;
}

/*
 * Abstract GEMV used by the solver loop so the same loop body serves the dense and the
 * generate-mode (structured) cases.
 */
typedef void (*oracle_matvec_fn)(const void *ctx, const double *p, double *Ap, size_t n);

static void matvec_dense(const void *ctx, const double *p, double *Ap, size_t n)
{
    oracle_gemv(1.0, (const double *)ctx, p, 0.0, Ap, n, n);
}

/*
 * Generate-mode GEMV without storing the n*n matrix.  For A = tridiag(1,2,1) stored dense the
 * reference's sequential row sum (OMP.hpp:255-259) adds exact zeros everywhere except columns
 * r-1, r, r+1, visited in that order starting from y_val = 0.0; beta = 0 and y = 0 on entry
 * (OMP.hpp:58, :70) so y[r] = 0*y[r] + y_val = y_val.  Hence
 *     y[r] = ((0 + 1*p[r-1]) + 2*p[r]) + 1*p[r+1]
 * with exactly those roundings — bit-identical to the dense evaluation (tests/test_oracle.py
 * checks that claim against oracle_gemv on the materialised matrix).  This is what lets the
 * oracle follow the n = 100 000 and n = 300 000 configs in O(n) memory.
 */
static void matvec_generated(const void *ctx, const double *p, double *Ap, size_t n)
{
    (void)ctx;
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < n; r++) {
        double y_val = 0.0;
        if (r > 0) y_val += 1.0 * 1.0 * p[r - 1];
        y_val += 1.0 * 2.0 * p[r];
        if (r + 1 < n) y_val += 1.0 * 1.0 * p[r + 1];
        Ap[r] = 0.0 * Ap[r] + y_val;
    }
}

/*
 * ref: CPU/ConjugateGradient_CPU_OMP.hpp:49-91 (solve), same loop as
 * CPU/ConjugateGradient_CPU_MPI_OMP.hpp:71-142.
 *   x0 = 0, r = p = b, rr = bb = b.b
 *   for it = 1..max_iters: Ap = A p; alpha = rr/(p.Ap); x += alpha p; r -= alpha Ap;
 *       rr_new = r.r; beta = rr_new/rr; rr = rr_new; if sqrt(rr/bb) < rel_error break;
 *       p = r + beta p
 * Returns 1 when converged (iters_out <= max_iters) else 0 with iters_out = max_iters + 1 — the
 * loop variable after exit, which is what the reference prints in its CSV (MPI_OMP.hpp:125).
 * rr_hist (nullable) receives sqrt(rr/bb) after every executed iteration, up to hist_cap.
 */
static int cg_loop(oracle_matvec_fn mv, const void *ctx, const double *b, double *x, size_t n,
                   int max_iters, double rel_error, int *iters_out, double *rel_out,
                   double *rr_hist, size_t hist_cap)
{
    double *r = (double *)malloc(n * sizeof(double));
    double *p = (double *)malloc(n * sizeof(double));
    double *Ap = (double *)malloc(n * sizeof(double));
    if (!r || !p || !Ap) { free(r); free(p); free(Ap); return -1; }

    for (size_t i = 0; i < n; i++) { Ap[i] = 0.0; x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }

    double alpha, beta, rr_new;
    double rhs_module = oracle_dot(b, b, n);
    double rr = rhs_module;
    int num_iters;
    for (num_iters = 1; num_iters <= max_iters; num_iters++) {
        mv(ctx, p, Ap, n);
        alpha = rr / oracle_dot(p, Ap, n);
        oracle_axpby(alpha, p, 1.0, x, n);
        oracle_axpby(-alpha, Ap, 1.0, r, n);
        rr_new = oracle_dot(r, r, n);
        beta = rr_new / rr;
        rr = rr_new;
        if (rr_hist && (size_t)(num_iters - 1) < hist_cap) rr_hist[num_iters - 1] = sqrt(rr / rhs_module);
        if (sqrt(rr / rhs_module) < rel_error) break;
        oracle_axpby(1.0, r, beta, p, n);
    }
    if (iters_out) *iters_out = num_iters;
    if (rel_out) *rel_out = sqrt(rr / rhs_module);
    free(r); free(p); free(Ap);
    return num_iters <= max_iters ? 1 : 0;
}

int oracle_cg_solve(const double *A, const double *b, double *x, size_t n, int max_iters,
                    double rel_error, int *iters_out, double *rel_out, double *rr_hist,
                    size_t hist_cap)
{
    return cg_loop(matvec_dense, A, b, x, n, max_iters, rel_error, iters_out, rel_out, rr_hist, hist_cap);
}

/* Generate mode end to end (test_CG_CPU_MPI_OMP.cpp:114-198: generate_matrix, generate_rhs,
 * solve) in O(n) memory; see matvec_generated. */
int oracle_cg_solve_generated(size_t n, double *x, int max_iters, double rel_error,
                              int *iters_out, double *rel_out, double *rr_hist, size_t hist_cap)
{
    double *b = (double *)malloc(n * sizeof(double));
    if (!b) return -1;
    oracle_generate_rhs(b, n);
    int rc = cg_loop(matvec_generated, NULL, b, x, n, max_iters, rel_error, iters_out, rel_out, rr_hist, hist_cap);
    free(b);
    return rc;
}

/* One structured GEMV, exported so tests can compare it bit-for-bit with oracle_gemv. */
void oracle_gemv_generated(const double *p, double *Ap, size_t n)
{
    for (size_t i = 0; i < n; i++) Ap[i] = 0.0;
    matvec_generated(NULL, p, Ap, n);
}

/* ref: challenge/main/random_spd_system.cpp:27-38 (random_matrix): glibc srand(seed) then
 * column-major fill with 2*rand()/RAND_MAX - 1.  Exposed so the numpy restatement of the
 * random-SPD *distribution* (oracle/random_spd.py) draws the very same stream. */
void oracle_rand_fill(double *out, size_t count, int seed)
{
    srand((unsigned)seed);
    for (size_t i = 0; i < count; i++) out[i] = ((2.0 * rand()) / RAND_MAX) - 1.0;
}

/* ---------------------------------------------------------------------------------------------
 * fp32 restatement: the same loop with every variable a float, i.e. what the reference's <float>
 * instantiations compute (template parameter FloatingType = float in OMP.hpp:49-91,219-263).
 * --------------------------------------------------------------------------------------------- */
static float dot_f32(const float *x, const float *y, size_t size)
{
    float result = 0.0f;
    for (size_t i = 0; i < size; i++) result += x[i] * y[i];
    return result;
}

static void axpby_f32(float alpha, const float *x, float beta, float *y, size_t size)
{
    for (size_t i = 0; i < size; i++) y[i] = alpha * x[i] + beta * y[i];
}

void oracle_gemv_f32(const float *A, const float *x, float *y, size_t rows, size_t cols)
{
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; r++) {
        float y_val = 0.0f;
        const float *a = A + r * cols;
        for (size_t c = 0; c < cols; c++) y_val += 1.0f * a[c] * x[c];
        y[r] = 0.0f * y[r] + y_val;
    }
}

/* A == NULL selects the generate-mode matrix (tridiag(1,2,1)) in structured form, as in matvec_generated. */
int oracle_cg_solve_f32(const float *A, const float *b, float *x, size_t n, int max_iters, float rel_error,
                        int *iters_out, float *rel_out)
{
    float *r = (float *)malloc(n * sizeof(float)), *p = (float *)malloc(n * sizeof(float)), *Ap = (float *)malloc(n * sizeof(float));
    if (!r || !p || !Ap) { free(r); free(p); free(Ap); return -1; }
    for (size_t i = 0; i < n; i++) { Ap[i] = 0.0f; x[i] = 0.0f; r[i] = b[i]; p[i] = b[i]; }
    float alpha, beta, rr_new;
    float rhs_module = dot_f32(b, b, n);
    float rr = rhs_module;
    int num_iters;
    for (num_iters = 1; num_iters <= max_iters; num_iters++) {
        if (A) {
            oracle_gemv_f32(A, p, Ap, n, n);
        } else {
            for (size_t q = 0; q < n; q++) {
                float y_val = 0.0f;
                if (q > 0) y_val += 1.0f * 1.0f * p[q - 1];
                y_val += 1.0f * 2.0f * p[q];
                if (q + 1 < n) y_val += 1.0f * 1.0f * p[q + 1];
                Ap[q] = 0.0f * Ap[q] + y_val;
            }
        }
        alpha = rr / dot_f32(p, Ap, n);
        axpby_f32(alpha, p, 1.0f, x, n);
        axpby_f32(-alpha, Ap, 1.0f, r, n);
        rr_new = dot_f32(r, r, n);
        beta = rr_new / rr;
        rr = rr_new;
        if (sqrtf(rr / rhs_module) < rel_error) break;
        axpby_f32(1.0f, r, beta, p, n);
    }
    if (iters_out) *iters_out = num_iters;
    if (rel_out) *rel_out = sqrtf(rr / rhs_module);
    free(r); free(p); free(Ap);
    return num_iters <= max_iters ? 1 : 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
