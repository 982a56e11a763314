/*
 * lamcg.h — C ABI of the B200-native dense Conjugate-Gradient library (liblamcg.so).
 *
 * This is the drop-in boundary for the CG path of edo01/2024-EUMaster4HPC-Student-Challenge.
 * The reference has no FFI: its boundary is the C++ class LAM::ConjugateGradient<T>
 * (challenge/main/LAM/src/ConjugateGradient.hpp:9-28) plus the generate-mode extensions of the
 * distributed classes (LAM/src/CPU/ConjugateGradient_CPU_MPI_OMP.hpp:31-35,
 * LAM/src/GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cuh:37-41).  Each entry point
 * below names the reference method it stands behind; the C++ class
 * LAM::ConjugateGradient_B200<T> (LAM/src/B200/ConjugateGradient_B200.hpp in this repo) and the
 * Python binding call nothing else.  Plain pointers and sizes only; every function returns
 * LAMCG_OK (0) or a negative lamcg_status, never throws, never exits.  There is no CPU fallback:
 * without a CUDA device every call that needs one fails with LAMCG_ERR_CUDA.
 *
 * Process model: one lamcg_t per GPU ("rank").  A single-GPU solve uses lamcg_create().  A
 * multi-GPU solve row-partitions A exactly like the reference (MPI_OMP.hpp:175-196: n/P rows per
 * rank, remainder to the last) with one rank per process (torchrun / the forking CLI) and either
 * NCCL collectives (lamcg_comm_init_nccl) or direct NVLink peer stores (lamcg_comm_init_peer).
 */
#ifndef LAMCG_H
#define LAMCG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lamcg lamcg_t;

typedef enum {
    LAMCG_OK = 0,
    LAMCG_ERR_INVALID = -1, /* bad argument / call order */
    LAMCG_ERR_CUDA = -2,    /* CUDA runtime error (message in lamcg_last_error) */
    LAMCG_ERR_IO = -3,      /* cannot open / short read / short write */
    LAMCG_ERR_SHAPE = -4,   /* matrix not square, rhs size mismatch, rhs cols != 1 */
    LAMCG_ERR_NOMEM = -5,   /* device or host allocation failed */
    LAMCG_ERR_COMM = -6,    /* NCCL / peer-exchange failure */
    LAMCG_ERR_STATE = -7,   /* e.g. solve before a matrix and rhs exist */
    LAMCG_ERR_DEVICE = -8   /* a kernel reported a fault (barrier timeout, non-finite scalar) */
} lamcg_status;

/* What ConjugateGradient::solve reports (OMP.hpp:80-90, MPI_OMP.hpp:122-141). */
typedef struct {
    int converged;         /* solve()'s bool: stop test sqrt(rr/bb) < rel_error met within max_iters */
    int iterations;        /* 1-based index of the stopping iteration; max_iters+1 when not converged
                              (the reference's loop variable after exit, MPI_OMP.hpp:125) */
    double rel_residual;   /* sqrt(rr/bb) at exit */
    double solve_seconds;  /* device time of the iteration loop (CUDA events) */
    double gemv_seconds;   /* summed device time of the timed GEMV launches (gemv_launches_timed of them); 0 unless option time_gemv */
    int iterations_run;    /* iterations actually executed on the device (== min(iterations,max_iters)) */
    int kernel_launches;   /* kernels of this library launched by this solve (incl. graph nodes) */
    int numerical_breakdown; /* 1: stopped early on a non-finite residual (b = 0, A not SPD): reported like the
                              reference would after max_iters NaN iterations: converged 0, max_iters+1, nan */
    int gemv_launches_timed; /* how many GEMV launches gemv_seconds sums over: every executed one with time_gemv = 1, one per
                              graph chunk with loop_mode 2 and time_gemv >= 2 */
} lamcg_result;

typedef struct {
    size_t n;           /* global system size (rows == cols) */
    size_t local_rows;  /* rows of A owned by this rank == ConjugateGradient_*::get_num_rows() */
    size_t row_offset;  /* first global row owned by this rank */
    size_t lda;         /* leading dimension (in elements) of the device row block, >= n */
    int rank, nranks, device, sm_count;
    int comm_mode;      /* 0 none (single rank), 1 NCCL, 2 peer stores */
    int has_matrix, has_rhs;
    int gemv_variant;   /* resolved K1 kernel: 36 / 32 row sweep with 128-bit loads (defaults: tall / short blocks), 46 / 42 the same
                           with 256-bit loads, 11 warp-per-rows with TMA-staged p, 2 TMA ring */
    int gemv_grid, gemv_block, gemv_smem_bytes;
    int dtype;          /* 0 fp64, 1 fp32 */
    int ingest_threads; /* reader threads and ... */
    int ingest_chunks;  /* ... staging chunks the last lamcg_load_matrix used on this rank */
    int matrix_elem_bytes; /* bytes per matrix element in HBM: 8 / 4 by dtype, 4 on an fp64 handle under option matrix_f32 */
    unsigned long long matrix_f32_inexact;  /* option matrix_f32: entries of the last matrix this rank took in whose fp32 value differs
                                               from the fp64 source (0: the solve runs on exactly the caller's matrix) ... */
    unsigned long long matrix_f32_overflow; /* ... and finite entries that became infinite */
} lamcg_info;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* One rank on CUDA device `device` (ordinal as seen by this process). */
int lamcg_create(lamcg_t **out, int device);
/* Rank `rank` of `nranks` of a row-partitioned job (MPI_Comm_rank/size in the reference). */
int lamcg_create_ranked(lamcg_t **out, int device, int rank, int nranks);
/* Same with the element type of everything STORED (A, b, x and the work vectors): dtype 0 = fp64 (what the
 * reference's drivers instantiate, and the type of lamcg_create / lamcg_create_ranked), 1 = fp32 (the <float>
 * instantiation of the reference classes, GPU/local/ConjugateGradient_MultiGPUS_CUDA.cu:539).  With fp32 storage
 * all reductions, alpha, beta and the stop test stay in fp64; files and caller buffers hold floats.  Every `void *`
 * data argument below points at elements of the handle's type. */
int lamcg_create_typed(lamcg_t **out, int device, int rank, int nranks, int dtype);
void lamcg_destroy(lamcg_t *h);
/* Last error message of this handle (or of the failed create when h == NULL). */
const char *lamcg_last_error(const lamcg_t *h);
const char *lamcg_version(void);

/* ---- options (all optional; also readable from env LAMCG_<KEY>) ----------------------------- */
/*  gemv_variant  0 auto | 36 / 32 row sweep, 128-bit loads | 46 / 42 row sweep, 256-bit loads | 11 warp rows + TMA-staged p | 2 TMA ring
 *  loop_mode     0 auto (single rank, fp64, n <= 16384: 3; else 2) | 1 stream | 2 graph | 3 persistent (one cooperative kernel)
 *  persist_variant  0 auto | 3 | 4: kernel of the one-kernel loop (3: K1's streaming sweep inside the loop, two scalar exchanges,
 *                auto from n = 4081 to 16384; 4: one all-gather of Ap per iteration with redundant scalars, n <= 4096, auto below)
 *  persist_poll_delay (default 650 cycles: a thread's first poll of the gathered Ap), persist_l2_keep_mb (default 64: megabytes of A
 *                the streaming kernel loads with the L2 evict-last policy), persist_rows_smem: tuning of the one-kernel loop
 *  persist_grid  upper bound on the CTAs of the persistent kernel (0: one per SM)
 *  fuse_updates  1 (default): K2 + K3 as one cooperative launch (single rank / peer mode); 0: two launches
 *  loop_profile  1: stream / graph loop in peer mode: CTA 0 accumulates wait and work cycles per phase (lamcg_get_loop_profile)
 *  matrix_f32    1 (fp64 handles; default 0, OPT-IN because it changes the problem being solved): the matrix block is held in HBM as
 *                fp32 while b, x, r, p, Ap, every product, sum and scalar stay fp64 (the reference instantiates <float>,
 *                GPU_MPI.cu:707; this is the mixed form).  K1 widens each element (exact) and then runs the fp64 sweep's unfused
 *                multiply + add, so the solve is the fp64 solve of fl32(A) at half the HBM bytes per iteration.  A matrix whose
 *                entries are fp32 numbers (generate mode: 0, 1, 2) is solved as given; otherwise lamcg_info.matrix_f32_inexact
 *                says how many entries were rounded.  Callers and files still hold doubles (narrowed on the device during
 *                set_matrix / load_matrix).  Changing the option drops the loaded system.  Row sweeps 32 / 36 and the stream /
 *                graph loops only; the one-kernel loop, the SPD generator and save_system need it off.
 *  spd_simt      1: the SPD generator as in round 1 (SIMT products, recursion to single columns); 0 (default): DMMA + CholeskyQR2 leaves
 *  chunk_iters   iterations per graph launch
 *  time_gemv     1: CUDA events around every GEMV launch (stream loop; with loop_mode 2 as event-record nodes inside the graph);
 *                2 with loop_mode 2: around one GEMV per graph chunk (an event node between two kernels is not free)
 *  ingest_threads (default 8), ingest_chunk_bytes (default 4 MB): reader threads / staging-chunk size of lamcg_load_matrix
 *  peer_timeout_s (default 600): bound of every in-kernel wait for a peer rank; on expiry the solve returns LAMCG_ERR_DEVICE
 *  gemv_ctas_per_sm  tuning override     history  0/1 keep sqrt(rr/bb) per iteration (default 1)
 *  debug_persist_fail  test hook: pretend the cooperative launch of the persistent kernel was refused */
int lamcg_set_option(lamcg_t *h, const char *key, long long value);
int lamcg_get_info(const lamcg_t *h, lamcg_info *out);

/* ---- multi-GPU bootstrap -------------------------------------------------------------------- */
/* NCCL: replaces ncclGetUniqueId + MPI_Bcast(id) + ncclCommInitRank of the reference
 * (GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:320-327).  Rank 0 fills a 128-byte id,
 * the host program broadcasts it (torch.distributed / shared memory), every rank calls init. */
#define LAMCG_NCCL_ID_BYTES 128
int lamcg_comm_nccl_unique_id(void *id_out);
int lamcg_comm_init_nccl(lamcg_t *h, const void *id);
/* Peer stores over NVLink (replaces the cudaMemcpyPeerAsync scatter/gather of
 * GPU/local/ConjugateGradient_MultiGPUS_CUDA.cu:336-376): every rank exports a handle to its
 * exchange buffer, the host all-gathers the handles, every rank imports all of them. */
#define LAMCG_PEER_HANDLE_BYTES 128
int lamcg_comm_peer_export(lamcg_t *h, size_t n, void *handle_out);
int lamcg_comm_init_peer(lamcg_t *h, const void *all_handles /* nranks * LAMCG_PEER_HANDLE_BYTES */);

/* ---- the system ----------------------------------------------------------------------------- */
/* generate_matrix(rows, cols) (MPI_OMP.hpp:167-256): this rank's rows of tridiag(1,2,1) stored
 * dense, written by a device kernel straight into HBM.  rows must equal cols. */
int lamcg_generate_matrix(lamcg_t *h, size_t rows, size_t cols);
/* generate_rhs() (MPI_OMP.hpp:144-165): b = 1. */
int lamcg_generate_rhs(lamcg_t *h);
/* load_matrix_from_file (OMP.hpp:137-197, MPI_OMP.hpp:307-417): 16-byte header (size_t rows,
 * size_t cols) + row-major doubles; this rank reads only its own row block, 64-bit sizes. */
int lamcg_load_matrix(lamcg_t *h, const char *path);
/* load_rhs_from_file (OMP.hpp:93-135): header cols must be 1 and rows must equal n. */
int lamcg_load_rhs(lamcg_t *h, const char *path);
/* In-memory system — the original challenge's solve(A, b, x, size, ...) signature
 * (test/test_CG_CPU_OMP.cpp:76-79).  A is row-major with leading dimension n and may be a host
 * or a device pointer.  layout 0: A is the whole n*n matrix (the rank takes its rows);
 * layout 1: A is only this rank's local_rows*n block. */
int lamcg_set_matrix(lamcg_t *h, const void *A, size_t n, int layout);
int lamcg_set_rhs(lamcg_t *h, const void *b, size_t n);

/* Random SPD system like challenge/main/random_spd_system.cpp (there: Intel MKL on the host):
 * Q = recursive block Gram-Schmidt of a U(-1,1) matrix drawn with glibc srand(seed)/rand(),
 * eigenvalues exp(3.5 U) (seed - 10), A = (Q sqrt(D))(Q sqrt(D))^T, rhs U(-1,1) (seed + 10).  The random
 * streams are drawn on the host, every O(n^3) step runs on the GPU.  Single rank. */
int lamcg_random_spd_system(lamcg_t *h, size_t n, int seed);
/* Write the current system in the reference's binary format (random_spd_system.cpp:105-121). Single rank. */
int lamcg_save_system(lamcg_t *h, const char *matrix_path, const char *rhs_path);

/* ---- solve ---------------------------------------------------------------------------------- */
/* solve(max_iters, rel_error) (OMP.hpp:49-91): x0 = 0, r = p = b; may be called repeatedly. */
int lamcg_solve(lamcg_t *h, int max_iters, double rel_error, lamcg_result *out);
/* Continue the last solve for up to `more_iters` further iterations, exactly as if it had been called with
 * max_iters + more_iters in the first place: x, r, p, rr and the iteration counter are still on the device, so
 * solve(k) followed by resume(m) is bit-identical to solve(k + m).  Needs a solve (stream or graph loop; set
 * loop_mode 2 for n <= 16384, where auto picks the one-kernel loop) that stopped on max_iters without converging, or a loaded checkpoint, and an unchanged
 * system; otherwise LAMCG_ERR_STATE.  `out` reports totals (iterations, iterations_run count from the original
 * start; solve_seconds is this call's loop time).  Collective over all ranks.  The reference has no equivalent: its
 * long generate-mode runs (n/2 iterations to converge, MPI_OMP.hpp:71-142) restart from x = 0. */
int lamcg_solve_resume(lamcg_t *h, int more_iters, double rel_error, lamcg_result *out);
/* Checkpoint of a solve that stopped on max_iters: this rank's slices of x, r and p, the scalars (bb, rr, beta), the
 * iteration count and the residual history, one file per rank (the caller names it).  lamcg_checkpoint_load puts that
 * state back into a handle that holds the same system with the same rank layout and element type (checked:
 * LAMCG_ERR_SHAPE); lamcg_solve_resume then carries on bit-identically to an uninterrupted solve. */
int lamcg_checkpoint_save(lamcg_t *h, const char *path);
int lamcg_checkpoint_load(lamcg_t *h, const char *path);
/* sqrt(rr/bb) after each executed iteration of the last solve; returns the count copied. */
int lamcg_get_residual_history(lamcg_t *h, double *out, int capacity);
/* This rank's slice of x (local_rows doubles, host pointer). */
int lamcg_get_solution_local(lamcg_t *h, void *x_local);
/* The whole x (n doubles, host pointer); collective over all ranks when nranks > 1. */
int lamcg_get_solution(lamcg_t *h, void *x);
/* save_result_to_file (OMP.hpp:199-217): header (n, 1) + x; collective, rank 0 writes.  Unlike
 * the reference the cols word is a clean 1 (SURVEY §2.4 defect 1) and x, not b, is written
 * (defect 2). */
int lamcg_save_solution(lamcg_t *h, const char *path);

/* ---- measurement / test hooks ---------------------------------------------------------------- */
/* One GEMV of the solver's own kernel on this rank's block: y_local = A_local * p, and the fused
 * epilogue value sum_i p[row_offset+i]*y_local[i].  Host pointers; p has n entries. */
int lamcg_gemv(lamcg_t *h, const void *p, void *y_local, double *p_dot_y);
/* K2 / K3 in isolation (single rank, no system needed; x, r, p, Ap are n-element host arrays of the handle's type, rr = r.r
 * entering the iteration, pAp = p.Ap): one pass of the vector kernels as they run inside the loop — fused != 0: the cooperative
 * update_fused_kernel, 0: update_xr_kernel then update_p_kernel.  On return x += alpha p, r -= alpha Ap, p = r + beta p
 * (OMP.hpp:72-78) and alpha = rr / pAp, rr_new = r.r, beta = rr_new / rr. */
int lamcg_vector_update_step(lamcg_t *h, size_t n, void *x, void *r, void *p, const void *Ap, double rr, double pAp, int fused,
                             double *alpha, double *rr_new, double *beta);
/* Launch the GEMV kernel `reps` times back to back on the solver's stream and return the average
 * device milliseconds per launch (CUDA events on that stream), after `warmup` untimed launches. */
int lamcg_time_gemv(lamcg_t *h, int warmup, int reps, double *ms_per_launch);
/* Persistent loop only: SM cycles CTA 0 spent in each phase of the last solve, summed over its iterations.
 * v3: [0] p update  [1] GEMV  [2] row sums + p.Ap exchange  [3] alpha broadcast  [4] x/r update + r.r exchange  [5] beta broadcast.
 * v4: [0] p update  [1] GEMV  [2] row sums + publish  [3] gather of Ap  [4] p.Ap, alpha, r, r.r  [5] beta, stop test.
 * Stream / graph loop with option loop_profile (peer mode): [2] wait for p.Ap  [3] x, r update + r.r  [4] wait for r.r
 * [5] beta, p update, peer stores, fence, flags.
 * Returns the count. */
int lamcg_get_loop_profile(lamcg_t *h, long long *cycles_out, int capacity);
/* Plain streaming read of this rank's block (sum of all elements): the read-only HBM ceiling the
 * GEMV is compared with.  Returns average ms per pass and the checksum. */
int lamcg_time_stream_read(lamcg_t *h, int warmup, int reps, double *ms_per_pass, double *checksum);

#ifdef __cplusplus
}
#endif
#endif /* LAMCG_H */
