// lamcg.cu — host side of liblamcg.so: the rank object, the device-resident CG loop (stream or
// CUDA-graph driven), multi-GPU plumbing, file ingest, and the extern "C" surface of
// include/lamcg.h.  No CPU fallback anywhere: every compute entry point needs the GPU.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <cerrno>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/lamcg.h"
#include "lamcg_kernels.cuh"
#include "lamcg_spd.cuh"
#include "nccl_dyn.h"

using namespace lamcgk;

namespace {

thread_local std::string g_create_error;

constexpr int kVecThreads = 256;
constexpr int kCommNone = 0, kCommNccl = 1, kCommPeer = 2;
constexpr int kLoopAuto = 0, kLoopStream = 1, kLoopGraph = 2, kLoopPersistent = 3;
constexpr size_t kPersistAutoMaxN = 16384;  // measured (profiles/r01_small_n_gen3.log): the one-kernel loop beats the graph loop up to here
constexpr size_t kPersistMaxN = 16384;      // p must fit in shared memory next to the task partials
constexpr int kPersistUnavailable = 1;      // solve_persistent: the cooperative launch was refused, nothing ran

struct GemvPlan {
    int variant = 0;
    void (*kernel)(GemvArgs) = nullptr;
    int grid = 0, block = 0;
    size_t smem = 0;
    int rows_per_pass = 1;
};

long long env_ll(const char *key, long long dflt)
{
    std::string name = "LAMCG_";
    for (const char *c = key; *c; ++c) name.push_back((char)toupper(*c));
    const char *v = getenv(name.c_str());
    return v && *v ? atoll(v) : dflt;
}

} // namespace

struct lamcg {
    int device = 0, rank = 0, nranks = 1, sm_count = 0;
    cudaStream_t stream = nullptr;
    size_t n = 0, local_rows = 0, row_offset = 0, lda = 0;
    size_t alloc_n = 0;
    int dtype = 0;                 // 0 = fp64 (the hot path), 1 = fp32 storage (reductions and scalars stay fp64)
    size_t esz = sizeof(double);   // bytes per stored element
    size_t asz = sizeof(double);   // bytes per stored MATRIX element: esz, or 4 under option matrix_f32 on an fp64 handle
    unsigned long long *narrow_stats = nullptr; // device [2]: inexact / overflowed entries of the last fp64 -> fp32 matrix ingest
    unsigned long long last_inexact = 0, last_overflow = 0;
    // byte pointers: element type is `dtype`
    char *A = nullptr, *b_full = nullptr, *x = nullptr, *r = nullptr, *Ap = nullptr, *p_full = nullptr;
    char *x_full = nullptr; // gather target for get_solution (multi-rank), [lda]
    double *partials = nullptr; // [2 * kMaxGrid]
    double *hist = nullptr;
    int hist_cap = 0;
    DevState *st = nullptr;
    DevState *h_st = nullptr; // pinned, 3 slots
    bool has_matrix = false, has_rhs = false;

    // options
    long long opt_gemv_variant = 0, opt_loop_mode = 0, opt_chunk_iters = 16, opt_time_gemv = 0, opt_history = 1;
    long long opt_gemv_ctas_per_sm = 0;
    long long opt_ingest_threads = 8;
    long long opt_ingest_chunk_bytes = 4ll << 20; // staging buffer size of the file ingest (tests shrink it to force many chunks)
    long long opt_peer_timeout_s = 600;   // bound of every peer flag wait; ranks may finish a cold-cache ingest minutes apart
    long long opt_persist_grid = 0;       // 0: one CTA per SM; k > 0: at most k CTAs in the one-kernel loop (tests: small-device behaviour)
    long long opt_debug_persist_fail = 0; // test hook: pretend the cooperative launch was refused
    long long opt_loop_profile = 0;       // multi-rank stream / graph loop: CTA 0 of K1 / K2+K3 accumulates wait and work cycles (lamcg_get_loop_profile)
    long long opt_fuse_updates = 1;       // K2 + K3 in one cooperative launch (single rank / peer mode)
    long long opt_spd_simt = 0;           // 1: the SPD generator's products on the SIMT kernel only (comparison / fallback)
    long long opt_matrix_f32 = 0;         // 1 (fp64 handles): the matrix is held in HBM as fp32, everything else stays fp64
    int clock_khz = 1965000;
    char *ingest_pool = nullptr;          // pinned staging buffers of the file ingest (kept between loads)
    size_t ingest_pool_bytes = 0;
    int last_ingest_threads = 0;
    long long last_ingest_chunks = 0;
    long long opt_persist_rows_smem = -1; // v4: -1: as many resident rows as fit; k >= 0: at most k
    long long opt_persist_variant = 0;    // 0 auto (v4 below lda = 4096, v3 from there) | 3 | 4 (n <= 4096)
    long long opt_persist_l2_keep_mb = 64; // generation 3: megabytes of A loaded evict-last (kept in L2 between iterations)
    long long opt_persist_poll_delay = 650; // v4: cycles before a thread's first poll of the gathered Ap

    // comm
    int comm_mode = kCommNone;
    ncclComm_t nccl = nullptr;
    // peer exchange (comm_mode == kCommPeer)
    unsigned char *peer_base = nullptr; // this rank's exchange buffer (cudaMalloc, IPC-exported)
    size_t peer_bytes = 0, peer_n = 0;
    PeerView pv{};
    unsigned long long seq_next = 1, gather_seq = 0;
    double *gemm_ws = nullptr; // split-K workspace of the SPD generator (alive only inside lamcg_random_spd_system)
    unsigned long long *persist_ll = nullptr; // [2][G][G][2] tagged partial words (per-CTA inboxes)
    size_t persist_ll_words = 0;
    std::vector<const void *> persist_warm; // persistent kernels that have had their first (slow, driver-side) launch

    // graph cache
    cudaGraphExec_t graph_exec = nullptr;
    // graph loop with per-GEMV timing (loop_mode 2 + time_gemv 1): two executables launched alternately, each with its own
    // externally recorded events around every K1, so that a chunk's events can be read while the next chunk runs
    cudaGraphExec_t graph_exec_timed[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> graph_events[2];
    int graph_chunk = 0;
    int graph_variant = 0;
    int graph_comm = -1;
    size_t graph_n = 0;

    GemvPlan plan;
    std::vector<cudaEvent_t> gemv_events;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_ring[2] = {nullptr, nullptr};
    int last_hist_count = 0;
    // resume / checkpoint: true while the device still holds the state of a solve that stopped on max_iters
    bool resumable = false;
    int done_iters = 0;                    // iterations that solve has executed
    unsigned long long cur_seq_base = 0;   // its peer sequence base
    std::string err;

    int fail(int code, const char *fmt, ...)
    {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return h->fail(LAMCG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

#define NCK(call)                                                                                     \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) return h->fail(LAMCG_ERR_COMM, "%s failed: %s", #call, nccl_api().GetErrorString(r_)); \
    } while (0)

namespace {

constexpr int kMaxGrid = 4096;
using lamcgk::kMaxRanks;

template <int RB, int CB, int ST>
bool plan_tma(lamcg *h, GemvPlan &p, int variant)
{
    using Cfg = GemvTmaCfg<RB, CB, ST>;
    p.variant = variant;
    p.kernel = lamcg_tmaring_kernel<RB, CB, ST>;
    p.block = Cfg::kThreads;
    p.smem = Cfg::kSmemBytes;
    p.rows_per_pass = RB;
    int per_sm = 1;
    p.grid = (int)std::min<size_t>((size_t)h->sm_count * per_sm, std::max<size_t>(1, h->local_rows));
    return cudaFuncSetAttribute(lamcg_tmaring_kernel<RB, CB, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess;
}

template <int U, int NT, int CPS, int VB>
bool plan_rowsweep(lamcg *h, GemvPlan &p, int variant)
{
    constexpr int R = 8;
    p.variant = variant;
    if (h->asz != h->esz) {
        if constexpr (VB == 16) p.kernel = lamcg_rowsweep_kernel<double, R, U, NT, CPS, VB, float>;
        else return false;
    } else if (h->dtype == 0) p.kernel = lamcg_rowsweep_kernel<double, R, U, NT, CPS, VB>;
    else p.kernel = lamcg_rowsweep_kernel<float, R, U, NT, CPS, VB>;
    p.block = NT;
    p.smem = 0;
    p.rows_per_pass = R;
    int per_sm = h->opt_gemv_ctas_per_sm > 0 ? (int)h->opt_gemv_ctas_per_sm : CPS;
    size_t want = (h->local_rows + R - 1) / R;
    p.grid = (int)std::min<size_t>((size_t)h->sm_count * per_sm, std::max<size_t>(1, want));
    return true;
}

template <int R, int U, int CPS = 2>
bool plan_warprows(lamcg *h, GemvPlan &p, int variant)
{
    p.variant = variant;
    p.kernel = lamcg_warprows_tmap_kernel<R, U, CPS>;
    p.block = kLdgWarps * 32;
    p.smem = kLdgSmemBytes;
    p.rows_per_pass = kLdgWarps * R;
    int per_sm = h->opt_gemv_ctas_per_sm > 0 ? (int)h->opt_gemv_ctas_per_sm : CPS;
    size_t want = (h->local_rows + p.rows_per_pass - 1) / p.rows_per_pass;
    p.grid = (int)std::min<size_t>((size_t)h->sm_count * per_sm, std::max<size_t>(1, want));
    return cudaFuncSetAttribute(lamcg_warprows_tmap_kernel<R, U, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess;
}

// K1 shapes kept after the round-1 sweeps (profiles/r01_sweep*.log; ~40 measured-and-lost shapes were deleted in round 2):
//   36  row sweep, 512 threads, 1 CTA/SM, 8 rows x 4 128-bit loads in flight   default for tall blocks (>= 12000 rows)
//   32  row sweep, 256 threads, 2 CTA/SM, same loads                           default for short blocks
//   46 / 42  the same two shapes with the sm_100 256-bit loads (8 rows x 2 loads of 32 bytes)
//   11  warp-per-rows, p staged in shared memory by TMA bulk copies (north_star's literal design; 7.1-7.2 TB/s)
//   2   whole A stream through a TMA bulk-copy ring (6.7-7.15 TB/s)
// 0 = auto.
int make_plan(lamcg *h)
{
    int v = (int)h->opt_gemv_variant;
    if (v == 0) v = h->local_rows >= 12000 ? 36 : 32;
    if (h->dtype != 0 && (v == 11 || v == 2))
        return h->fail(LAMCG_ERR_INVALID, "gemv_variant %d is fp64 only (fp32 handles use the row-sweep family 32/36/42/46)", v);
    if (h->asz != h->esz && v != 32 && v != 36)
        return h->fail(LAMCG_ERR_INVALID, "gemv_variant %d does not read an fp32 matrix (option matrix_f32 runs the row sweeps 32 / 36)", v);
    bool ok = false;
    GemvPlan p;
    switch (v) {
    case 32: ok = plan_rowsweep<4, 256, 2, 16>(h, p, v); break;
    case 36: ok = plan_rowsweep<4, 512, 1, 16>(h, p, v); break;
    case 42: ok = plan_rowsweep<2, 256, 2, 32>(h, p, v); break;
    case 46: ok = plan_rowsweep<2, 512, 1, 32>(h, p, v); break;
    case 11: ok = plan_warprows<4, 4>(h, p, v); break;
    case 2: ok = plan_tma<16, 256, 6>(h, p, v); break;
    default: return h->fail(LAMCG_ERR_INVALID, "unknown gemv_variant %d (32, 36, 42, 46, 11, 2)", v);
    }
    if (!ok) return h->fail(LAMCG_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed for gemv variant %d: %s", v,
                            cudaGetErrorString(cudaGetLastError()));
    if (p.grid > kMaxGrid) p.grid = kMaxGrid;
    h->plan = p;
    return LAMCG_OK;
}

void destroy_graphs(lamcg *h)
{
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    for (int k = 0; k < 2; ++k)
        if (h->graph_exec_timed[k]) { cudaGraphExecDestroy(h->graph_exec_timed[k]); h->graph_exec_timed[k] = nullptr; }
}

// Unmap every peer exchange buffer this rank has opened (also the partial set of a failed lamcg_comm_init_peer).
void close_peer_handles(lamcg *h)
{
    for (int r = 0; r < kMaxRanks; ++r) {
        if (r != h->rank && h->pv.base[r]) cudaIpcCloseMemHandle(h->pv.base[r]);
        h->pv.base[r] = nullptr;
    }
}

long long peer_timeout_cycles(const lamcg *h)
{
    const long long s = std::max<long long>(1, std::min<long long>(h->opt_peer_timeout_s, 86400));
    return s * (long long)h->clock_khz * 1000ll;
}

void free_system(lamcg *h)
{
    cudaSetDevice(h->device);
    destroy_graphs(h);
    cudaFree(h->A); cudaFree(h->b_full); cudaFree(h->x); cudaFree(h->r); cudaFree(h->Ap);
    cudaFree(h->p_full); cudaFree(h->x_full);
    h->A = h->b_full = h->x = h->r = h->Ap = h->p_full = h->x_full = nullptr;
    h->alloc_n = 0;
    h->has_matrix = h->has_rhs = false;
}

// Row partition of the reference (MPI_OMP.hpp:175-184): n/P rows each, remainder to the last rank.
void partition(size_t n, int nranks, int rank, size_t *rows, size_t *offset)
{
    const size_t base = n / (size_t)nranks;
    *offset = base * (size_t)rank;
    *rows = base + (rank == nranks - 1 ? n % (size_t)nranks : 0);
}

int alloc_system(lamcg *h, size_t n)
{
    if (n == 0) return h->fail(LAMCG_ERR_SHAPE, "empty system");
    if (h->alloc_n == n && h->A) return LAMCG_OK;
    if (h->comm_mode == kCommPeer && n != h->peer_n)
        return h->fail(LAMCG_ERR_SHAPE, "system size %zu differs from the size %zu the peer exchange buffers were exported for", n, h->peer_n);
    free_system(h);
    CK(cudaSetDevice(h->device));
    h->n = n;
    partition(n, h->nranks, h->rank, &h->local_rows, &h->row_offset);
    h->lda = (n + 15) / 16 * 16; // rows start on 128-byte boundaries; pad columns are zero
    const size_t rows_alloc = std::max<size_t>(h->local_rows, 1);
    cudaError_t e = cudaMalloc(&h->A, rows_alloc * h->lda * h->asz);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return h->fail(LAMCG_ERR_NOMEM, "cudaMalloc of the %zu x %zu row block (%.2f GB) failed: %s", h->local_rows, h->lda,
                       rows_alloc * h->lda * (double)h->asz / 1e9, cudaGetErrorString(e));
    }
    CK(cudaMalloc(&h->b_full, h->lda * h->esz));
    CK(cudaMalloc(&h->p_full, h->lda * h->esz));
    CK(cudaMalloc(&h->x_full, h->lda * h->esz));
    CK(cudaMalloc(&h->x, rows_alloc * h->esz));
    CK(cudaMalloc(&h->r, rows_alloc * h->esz));
    CK(cudaMalloc(&h->Ap, rows_alloc * h->esz));
    CK(cudaMemsetAsync(h->b_full, 0, h->lda * h->esz, h->stream));
    CK(cudaMemsetAsync(h->p_full, 0, h->lda * h->esz, h->stream));
    CK(cudaMemsetAsync(h->x, 0, rows_alloc * h->esz, h->stream));
    h->alloc_n = n;
    int rc = make_plan(h);
    if (rc != LAMCG_OK) return rc;
    return LAMCG_OK;
}

// ---- option matrix_f32: fp64 sources are narrowed on the device into the fp32 matrix block --------------------------------
bool mixed_storage(const lamcg *h) { return h->asz != h->esz; }

int narrow_begin(lamcg *h)
{
    if (!h->narrow_stats) CK(cudaMalloc(&h->narrow_stats, 2 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(h->narrow_stats, 0, 2 * sizeof(unsigned long long), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LAMCG_OK;
}

// rows [first_row, first_row + nr) of the local block from nr fp64 rows in device memory (row pitch src_ld), on stream st
cudaError_t narrow_rows(lamcg *h, const double *dev_src, size_t src_ld, size_t first_row, size_t nr, cudaStream_t st)
{
    if (nr == 0) return cudaSuccess;
    const long long total = (long long)nr * (long long)((h->n + 1) / 2);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 8);
    narrow_rows_kernel<<<grid, 256, 0, st>>>(dev_src, (long long)src_ld, reinterpret_cast<float *>(h->A) + first_row * h->lda, (long long)h->lda,
                                             (long long)nr, (long long)h->n, h->narrow_stats);
    return cudaGetLastError();
}

int narrow_end(lamcg *h)
{
    unsigned long long st[2] = {0, 0};
    CK(cudaMemcpy(st, h->narrow_stats, sizeof st, cudaMemcpyDeviceToHost));
    h->last_inexact = st[0];
    h->last_overflow = st[1];
    return LAMCG_OK;
}

char *p_ptr(lamcg *h, int par)
{
    if (h->comm_mode == kCommPeer) return reinterpret_cast<char *>(h->peer_base + h->pv.off_p[par & 1]);
    return h->p_full;
}

PeerView no_peer()
{
    PeerView v{};
    v.nranks = 0;
    return v;
}

GemvArgs gemv_args(lamcg *h, int check_done, int par)
{
    GemvArgs g;
    g.A = h->A;
    g.p = p_ptr(h, par);
    g.par = par;
    g.pv = (h->comm_mode == kCommPeer && check_done) ? h->pv : no_peer();
    g.Ap = h->Ap;
    g.partials = h->partials;
    g.st = h->st;
    g.rows = (long long)h->local_rows;
    g.lda = (long long)h->lda;
    g.row_offset = (long long)h->row_offset;
    g.check_done = check_done;
    return g;
}

int launch_gemv(lamcg *h, int check_done, int par = 0)
{
    GemvArgs g = gemv_args(h, check_done, par);
    h->plan.kernel<<<h->plan.grid, h->plan.block, h->plan.smem, h->stream>>>(g);
    CK(cudaGetLastError());
    return LAMCG_OK;
}

int vec_grid(lamcg *h)
{
    size_t want = (h->local_rows + kVecThreads - 1) / kVecThreads;
    return (int)std::min<size_t>(std::max<size_t>(want, 1), (size_t)h->sm_count * 4); // 4 CTAs of 256 threads per SM: always co-resident
}

// K2 and K3 as one cooperative launch unless NCCL has to run between them (or option fuse_updates = 0)
bool fuse_updates(const lamcg *h) { return h->comm_mode != kCommNccl && h->opt_fuse_updates != 0; }

VecArgs vec_args(lamcg *h, int par)
{
    VecArgs v;
    v.st = h->st;
    const bool multi = h->comm_mode == kCommNccl;
    v.pAp_src = multi ? &h->st->pAp : &h->st->pAp_local;
    v.rrn_src = multi ? &h->st->rrn : &h->st->rrn_local;
    v.x = h->x;
    v.r = h->r;
    v.Ap = h->Ap;
    v.p_in = p_ptr(h, par);
    v.p_out = p_ptr(h, par ^ 1);
    v.pv = h->comm_mode == kCommPeer ? h->pv : no_peer();
    v.partials = h->partials + kMaxGrid;
    v.hist = h->opt_history ? h->hist : nullptr;
    v.rows = (long long)h->local_rows;
    v.row_offset = (long long)h->row_offset;
    v.par = par;
    v.fused = fuse_updates(h) ? 1 : 0;
    v.prof = h->opt_loop_profile ? 1 : 0;
    return v;
}

// All-gather of the p slices with the reference partition: P equal slices of n/P plus the
// remainder owned by the last rank (MPI_OMP.hpp:505 gathers Ap the same way with Allgatherv).
int allgather_vec(lamcg *h, char *full)
{
    NcclApi &N = nccl_api();
    const ncclDataType_t dt = h->dtype == 0 ? ncclDouble : ncclFloat;
    const size_t base = h->n / (size_t)h->nranks;
    const size_t tail = h->n % (size_t)h->nranks;
    if (base > 0) NCK(N.AllGather(full + h->row_offset * h->esz, full, base, dt, h->nccl, h->stream));
    if (tail > 0) {
        char *t = full + base * (size_t)h->nranks * h->esz;
        NCK(N.Broadcast(t, t, tail, dt, h->nranks - 1, h->nccl, h->stream));
    }
    return LAMCG_OK;
}

// One CG iteration enqueued on h->stream (also the body captured into the CUDA graph).
int enqueue_iteration(lamcg *h, int par, cudaEvent_t ev0, cudaEvent_t ev1, int *launches, bool capturing = false)
{
    NcclApi &N = nccl_api();
    // inside a stream capture an event record must be flagged external to become a record NODE (a timestamp at every replay)
    const unsigned int evflags = capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
    if (ev0) CK(cudaEventRecordWithFlags(ev0, h->stream, evflags));
    int rc = launch_gemv(h, 1, par);
    if (rc != LAMCG_OK) return rc;
    if (ev1) CK(cudaEventRecordWithFlags(ev1, h->stream, evflags));
    if (h->comm_mode == kCommNccl)
        NCK(N.AllReduce(&h->st->pAp_local, &h->st->pAp, 1, ncclDouble, ncclSum, h->nccl, h->stream));
    VecArgs v = vec_args(h, par);
    const int vg = vec_grid(h);
    if (v.fused) {
        void *params[] = {&v};
        const void *fn = h->dtype == 0 ? (const void *)update_fused_kernel<double> : (const void *)update_fused_kernel<float>;
        CK(cudaLaunchCooperativeKernel(fn, dim3(vg), dim3(kVecThreads), params, 0, h->stream));
        *launches += 2;
        return LAMCG_OK;
    }
    if (h->dtype == 0) update_xr_kernel<double><<<vg, kVecThreads, 0, h->stream>>>(v);
    else update_xr_kernel<float><<<vg, kVecThreads, 0, h->stream>>>(v);
    CK(cudaGetLastError());
    if (h->comm_mode == kCommNccl)
        NCK(N.AllReduce(&h->st->rrn_local, &h->st->rrn, 1, ncclDouble, ncclSum, h->nccl, h->stream));
    if (h->dtype == 0) update_p_kernel<double><<<vg, kVecThreads, 0, h->stream>>>(v);
    else update_p_kernel<float><<<vg, kVecThreads, 0, h->stream>>>(v);
    CK(cudaGetLastError());
    if (h->comm_mode == kCommNccl) {
        rc = allgather_vec(h, h->p_full);
        if (rc != LAMCG_OK) return rc;
    }
    *launches += 3;
    return LAMCG_OK;
}

// Capture `chunk` iterations into one executable graph; timed_slot >= 0: with external event records around every K1
int capture_chunk(lamcg *h, int chunk, int timed_slot, cudaGraphExec_t *exec_out)
{
    if (timed_slot >= 0) {
        std::vector<cudaEvent_t> &ev = h->graph_events[timed_slot];
        while ((int)ev.size() < 2 * chunk) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            ev.push_back(e);
        }
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int dummy = 0;
    int rc = LAMCG_OK;
    for (int i = 0; i < chunk && rc == LAMCG_OK; ++i) {
        // time_gemv 1: events around every K1 of the chunk; time_gemv >= 2: around ONE K1 in the middle of the chunk (an event-record
        // node between two kernels costs ~7 us on a busy 8-GPU box: 14 us per iteration with events around every K1, measured)
        const bool timed = timed_slot >= 0 && (h->opt_time_gemv == 1 || i == chunk / 2);
        cudaEvent_t e0 = timed ? h->graph_events[timed_slot][2 * i] : nullptr;
        cudaEvent_t e1 = timed ? h->graph_events[timed_slot][2 * i + 1] : nullptr;
        rc = enqueue_iteration(h, i & 1, e0, e1, &dummy, true);
    }
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc != LAMCG_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return h->fail(LAMCG_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(exec_out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return h->fail(LAMCG_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return LAMCG_OK;
}

int build_graph(lamcg *h, int chunk, bool timed)
{
    const bool same = h->graph_chunk == chunk && h->graph_variant == h->plan.variant && h->graph_n == h->n && h->graph_comm == h->comm_mode;
    if (!same) destroy_graphs(h);
    int rc = LAMCG_OK;
    if (timed) {
        for (int k = 0; k < 2 && rc == LAMCG_OK; ++k)
            if (!h->graph_exec_timed[k]) rc = capture_chunk(h, chunk, k, &h->graph_exec_timed[k]);
    } else if (!h->graph_exec) {
        rc = capture_chunk(h, chunk, -1, &h->graph_exec);
    }
    if (rc != LAMCG_OK) return rc;
    h->graph_chunk = chunk;
    h->graph_variant = h->plan.variant;
    h->graph_comm = h->comm_mode;
    h->graph_n = h->n;
    return LAMCG_OK;
}

int ensure_hist(lamcg *h, int max_iters, int keep)
{
    if (!h->opt_history) return LAMCG_OK;
    int want = std::max(max_iters, 1);
    if (want > (1 << 22)) want = 1 << 22;
    if (h->hist_cap >= want) return LAMCG_OK;
    double *grown = nullptr;
    CK(cudaMalloc(&grown, (size_t)want * sizeof(double)));
    if (keep > 0 && h->hist) // a resumed solve keeps the history of the iterations already done
        CK(cudaMemcpyAsync(grown, h->hist, (size_t)std::min(keep, h->hist_cap) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->hist);
    h->hist = grown;
    h->hist_cap = want;
    destroy_graphs(h); // hist pointer is baked in
    return LAMCG_OK;
}

int check_device_error(lamcg *h, const DevState &s)
{
    if (s.error != 0) return h->fail(LAMCG_ERR_DEVICE, "device reported fault %d (1 = mbarrier timeout, 2 = peer flag timeout: a peer rank did not reach the exchange within peer_timeout_s, 3 = exchange timeout inside the one-kernel loop)", s.error);
    return LAMCG_OK;
}

// ---- persistent single-kernel loop (loop_mode 3 / auto for small single-rank systems) ------------
int solve_persistent(lamcg *h, int max_iters, double rel_error, lamcg_result *out)
{
    if (h->nranks != 1) return h->fail(LAMCG_ERR_INVALID, "the persistent loop is single-rank");
    if (h->dtype != 0) return h->fail(LAMCG_ERR_INVALID, "the persistent loop is fp64 only");
    if (mixed_storage(h)) return h->fail(LAMCG_ERR_INVALID, "the persistent loop reads an fp64 matrix (option matrix_f32 is on)");
    if (h->n > kPersistMaxN) return h->fail(LAMCG_ERR_INVALID, "the persistent loop supports n <= %zu", kPersistMaxN);

    int grid = (int)std::min<size_t>(std::min<size_t>((size_t)h->sm_count, (size_t)kPersistMaxGrid), h->n);
    if (h->opt_persist_grid > 0) grid = (int)std::min<long long>(grid, h->opt_persist_grid);
    // v3: [2][G][G][kLLStride] tagged scalar words; v4: [2][lda][2] tagged entries of the gathered Ap
    const size_t ll_words = std::max((size_t)2 * kLLStride * grid * grid, (size_t)4 * h->lda);
    if (!h->persist_ll || h->persist_ll_words < ll_words) {
        cudaFree(h->persist_ll);
        h->persist_ll = nullptr;
        CK(cudaMalloc(&h->persist_ll, ll_words * sizeof(unsigned long long)));
        h->persist_ll_words = ll_words;
    }
    const int rows_max = (int)((h->n + grid - 1) / grid);
    // both kernels handle one owned row per thread (x, r, Ap of row r0 + tid): a device that offers few SMs (MIG slice, MPS limit)
    // cannot run n near 16384 in one kernel -> the caller takes the graph loop instead
    if (rows_max > kPersistThreads) {
        h->fail(LAMCG_ERR_INVALID, "the one-kernel CG loop needs ceil(n / CTAs) <= %d rows per CTA (n = %zu on %d CTAs gives %d)", kPersistThreads, h->n, grid, rows_max);
        return kPersistUnavailable;
    }
    int dev_smem_max = 0;
    CK(cudaDeviceGetAttribute(&dev_smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    // persist_variant: 0 auto | 3 streaming sweep inside the loop (auto from lda = 4096) | 4 one exchange per iteration, p in
    // registers (n <= 4096; auto below lda = 4096).  Measured (profiles/r02_small_n_gen4_final.log, r02_gen3_l2keep_sweep.log):
    // n = 2048 237 k it/s (round 1's second generation: 145 k), n = 3000 102 k (first generation: 87 k); at n = 4096 the sweep's
    // 45 k beats 36 k.
    if (h->opt_persist_variant != 0 && h->opt_persist_variant != 3 && h->opt_persist_variant != 4)
        return h->fail(LAMCG_ERR_INVALID, "persist_variant must be 0 (auto), 3 or 4");
    if (h->opt_persist_variant == 4 && h->lda > 4096) return h->fail(LAMCG_ERR_INVALID, "persist_variant 4 needs n <= 4096");
    const bool v4 = h->opt_persist_variant == 4 || (h->opt_persist_variant == 0 && h->lda < 4096);
    const void *kernel;
    size_t fixed;
    if (v4) {
        kernel = h->lda <= 1024 ? (const void *)cg_persistent_v4_kernel<2> : h->lda <= 2048 ? (const void *)cg_persistent_v4_kernel<4>
                                                                                           : (const void *)cg_persistent_v4_kernel<8>;
        const size_t rows_pad = ((size_t)rows_max + 7) & ~(size_t)7;
        fixed = rows_pad * (kPersistThreads / 32) * sizeof(double); // per-warp row partials
    } else {
        kernel = (const void *)cg_persistent_v3_kernel<8, 2>;
        fixed = (h->lda + (((size_t)rows_max + 7) & ~(size_t)7)) * sizeof(double); // p | Ap of the CTA's rows
    }
    cudaFuncAttributes fattr;
    CK(cudaFuncGetAttributes(&fattr, kernel));
    const size_t stat = fattr.sharedSizeBytes; // static shared memory counts against the same per-block limit
    const size_t budget = (size_t)dev_smem_max > fixed + stat ? (size_t)dev_smem_max - fixed - stat : 0;
    int rows_smem = v4 ? (int)std::min<size_t>((size_t)rows_max, budget / (h->lda * sizeof(double))) : 0;
    if (rows_smem < rows_max && rows_smem > 8) rows_smem &= ~7; // whole 8-row groups from one place: the straight-line GEMV path
    if (h->opt_persist_rows_smem >= 0) rows_smem = std::min(rows_smem, (int)h->opt_persist_rows_smem);
    const size_t smem = fixed + (size_t)rows_smem * h->lda * sizeof(double);
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int max_blocks = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, kernel, kPersistThreads, smem));
    if (max_blocks < 1) return h->fail(LAMCG_ERR_CUDA, "persistent kernel does not fit on an SM (smem %zu)", smem);

    PersistArgs a;
    a.A = reinterpret_cast<const double *>(h->A);
    a.b = reinterpret_cast<const double *>(h->b_full);
    a.x = reinterpret_cast<double *>(h->x);
    a.r = reinterpret_cast<double *>(h->r);
    a.hist = h->opt_history ? h->hist : nullptr;
    a.ll = h->persist_ll;
    a.st = h->st;
    a.n = (long long)h->n;
    a.lda = (long long)h->lda;
    a.eps = rel_error;
    a.max_iters = max_iters;
    a.hist_cap = h->opt_history ? h->hist_cap : 0;
    a.rows_smem = rows_smem;
    a.rows_max = rows_max;
    a.poll_delay = (int)std::max<long long>(0, std::min<long long>(h->opt_persist_poll_delay, 100000));
    {   // v3: rows per CTA to keep in L2 between iterations (option persist_l2_keep_mb; measured, profiles/r02_gen3_l2keep_sweep.log:
        // n = 4096 42.6 -> 45.1 k it/s, 5000 26.2 -> 28.0 k, 8192 12.6 -> 12.8 k, 10000 8.39 -> 8.53 k with 64 MB; more than ~90 MB loses)
        const size_t keep_bytes = (size_t)std::max<long long>(0, std::min<long long>(h->opt_persist_l2_keep_mb, 120)) << 20;
        a.rows_l2keep = (int)(keep_bytes / std::max<size_t>(1, (size_t)grid * h->lda * sizeof(double)));
    }
    void *params[] = {&a};
    // The first cooperative launch of a kernel spends tens of milliseconds in the driver (module load, cooperative-launch setup).
    // Taken once per kernel OUTSIDE the timed region with a zero-iteration launch, so that solve_seconds of a first solve is
    // the loop and not the driver (the smoke test printed 22 k it/s for a 4 ms solve).
    if (!h->opt_debug_persist_fail && std::find(h->persist_warm.begin(), h->persist_warm.end(), kernel) == h->persist_warm.end()) {
        PersistArgs w = a;
        w.max_iters = 0;
        w.hist = nullptr;
        void *wparams[] = {&w};
        CK(cudaMemsetAsync(h->st, 0, sizeof(DevState), h->stream));
        if (cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kPersistThreads), wparams, smem, h->stream) == cudaSuccess) {
            CK(cudaStreamSynchronize(h->stream));
            h->persist_warm.push_back(kernel);
        } else {
            cudaGetLastError(); // the real launch below reports (or falls back)
        }
    }
    CK(cudaMemsetAsync(h->persist_ll, 0, ll_words * sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->st, 0, sizeof(DevState), h->stream));
    CK(cudaEventRecord(h->ev_start, h->stream));
    // A cooperative launch needs every CTA resident at once; when the device cannot grant that (SMs held by another context, MPS
    // limits) the caller falls back to the graph loop instead of failing the solve.  Option debug_persist_fail simulates it (test hook).
    cudaError_t le = h->opt_debug_persist_fail ? cudaErrorCooperativeLaunchTooLarge
                                               : cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kPersistThreads), params, smem, h->stream);
    if (le != cudaSuccess) {
        cudaGetLastError();
        h->fail(LAMCG_ERR_CUDA, "cooperative launch of the one-kernel CG loop failed: %s", cudaGetErrorString(le));
        return kPersistUnavailable;
    }
    CK(cudaEventRecord(h->ev_stop, h->stream));
    CK(cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "the persistent CG kernel faulted: %s", cudaGetErrorString(se));
    const DevState &s = h->h_st[2];
    int rc = check_device_error(h, s);
    if (rc != LAMCG_OK) return rc;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
    h->last_hist_count = h->opt_history ? std::min(s.iters_done, h->hist_cap) : 0;
    if (out) {
        out->converged = s.converged;
        out->iterations = s.converged ? s.iters_done : (max_iters < 0 ? 1 : max_iters + 1);
        out->rel_residual = std::sqrt(s.rr_final / s.bb);
        out->solve_seconds = ms * 1e-3;
        out->gemv_seconds = 0.0;
        out->iterations_run = s.iters_done;
        out->kernel_launches = 1;
        out->numerical_breakdown = s.breakdown;
        out->gemv_launches_timed = 0;
    }
    return LAMCG_OK;
}

// ---- file helpers ------------------------------------------------------------------------------
constexpr int kIngestSlots = 3;        // staging buffers per reader thread
constexpr int kIngestMaxThreads = 32;

int ensure_ingest_pool(lamcg *h, size_t bytes)
{
    if (h->ingest_pool && h->ingest_pool_bytes >= bytes) return LAMCG_OK;
    if (h->ingest_pool) cudaFreeHost(h->ingest_pool);
    h->ingest_pool = nullptr;
    h->ingest_pool_bytes = 0;
    cudaError_t e = cudaMallocHost(&h->ingest_pool, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        h->ingest_pool = nullptr;
        return h->fail(LAMCG_ERR_NOMEM, "cudaMallocHost of %.1f MB of ingest staging buffers failed: %s", bytes / 1e6, cudaGetErrorString(e));
    }
    h->ingest_pool_bytes = bytes;
    return LAMCG_OK;
}

int read_header(lamcg *h, int fd, const char *path, size_t *rows, size_t *cols)
{
    uint64_t hdr[2];
    ssize_t got = pread(fd, hdr, sizeof hdr, 0);
    if (got != (ssize_t)sizeof hdr) return h->fail(LAMCG_ERR_IO, "%s: short read of the 16-byte header", path);
    *rows = (size_t)hdr[0];
    *cols = (size_t)hdr[1];
    return LAMCG_OK;
}

int pread_full(int fd, void *buf, size_t bytes, off_t off)
{
    char *c = static_cast<char *>(buf);
    while (bytes > 0) {
        ssize_t got = pread(fd, c, bytes, off);
        if (got < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        if (got == 0) return -1;
        c += got;
        off += got;
        bytes -= (size_t)got;
    }
    return 0;
}

} // namespace

namespace {

int resolve_loop_mode(lamcg *h)
{
    int loop_mode = (int)h->opt_loop_mode;
    if (loop_mode == kLoopAuto) {
        if (h->opt_time_gemv) loop_mode = kLoopStream;
        else if (h->nranks == 1 && h->dtype == 0 && !mixed_storage(h) && h->n <= kPersistAutoMaxN) loop_mode = kLoopPersistent;
        else loop_mode = kLoopGraph;
    }
    // per-GEMV timing: stream launches with events around every K1, or (loop_mode 2 asked for explicitly) the graph loop with
    // externally recorded events inside the captured chunks; the one-kernel loop has no GEMV launches to time
    if (h->opt_time_gemv && loop_mode != kLoopGraph) loop_mode = kLoopStream;
    return loop_mode;
}

// Enqueue iterations first_iter .. max_total-1 (0-based) on the stream / as graph launches, watch the device's `done`
// latch one chunk behind, and report.  The device state (x, r, p, scalars, counters) must already be in place:
// init_solve_kernel for a fresh solve, resume_kernel for a continued one.  `launches` counts what the caller enqueued.
int run_loop(lamcg *h, int loop_mode, int first_iter, int max_total, lamcg_result *out, int launches)
{
    int chunk = (int)std::max<long long>(2, h->opt_chunk_iters);
    chunk += chunk & 1; // the parity double-buffering needs an even number of iterations per chunk
    int rc;
    const bool time_gemv = h->opt_time_gemv != 0 && max_total > first_iter;
    const bool graph_timed = time_gemv && loop_mode == kLoopGraph;
    if (loop_mode == kLoopGraph) {
        rc = build_graph(h, chunk, graph_timed);
        if (rc != LAMCG_OK) return rc;
    }
    std::vector<float> graph_gemv_ms; // graph_timed: per-iteration K1 durations in launch order (-1: launch not timed)
    auto read_graph_events = [&](int slot) { // the chunk launched from executable `slot` has completed
        for (int i = 0; i < chunk; ++i) {
            float t = -1.f;
            if (h->opt_time_gemv == 1 || i == chunk / 2)
                if (cudaEventElapsedTime(&t, h->graph_events[slot][2 * i], h->graph_events[slot][2 * i + 1]) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
            graph_gemv_ms.push_back(t);
        }
    };
    std::vector<int> chunk_slot; // per enqueued chunk c: timed executable it was launched from, or -1
    if (time_gemv && !graph_timed) {
        const size_t need = 2 * (size_t)std::min(max_total - first_iter, 1 << 16);
        while (h->gemv_events.size() < need) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            h->gemv_events.push_back(e);
        }
    }
    CK(cudaEventRecord(h->ev_start, h->stream));
    int launched = first_iter, c = 0, timed_iters = 0;
    bool stop = false;
    while (launched < max_total && !stop) {
        if (loop_mode == kLoopGraph && (launched & 1) == 0) {
            const int slot = graph_timed ? (int)(chunk_slot.size() & 1) : -1;
            CK(cudaGraphLaunch(graph_timed ? h->graph_exec_timed[slot] : h->graph_exec, h->stream)); // the captured chunk starts on parity 0
            chunk_slot.push_back(slot);
            launched += chunk;
            launches += (fuse_updates(h) ? 2 : 3) * chunk;
        } else {
            chunk_slot.push_back(-1);
            const int count = loop_mode == kLoopGraph ? 1 : chunk; // graph loop resumed on an odd iteration: one plain step first
            for (int i = 0; i < count && launched < max_total; ++i, ++launched) {
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                const size_t k = (size_t)(launched - first_iter);
                if (time_gemv && !graph_timed && 2 * k + 1 < h->gemv_events.size()) {
                    e0 = h->gemv_events[2 * k];
                    e1 = h->gemv_events[2 * k + 1];
                    timed_iters = (int)k + 1;
                }
                rc = enqueue_iteration(h, launched & 1, e0, e1, &launches);
                if (rc != LAMCG_OK) return rc;
            }
        }
        CK(cudaMemcpyAsync(&h->h_st[c & 1], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev_ring[c & 1], h->stream));
        if (c >= 1) { // one chunk of look-ahead: inspect the chunk before the one just enqueued
            CK(cudaEventSynchronize(h->ev_ring[(c - 1) & 1]));
            const DevState &s = h->h_st[(c - 1) & 1];
            if (s.done || s.error) stop = true;
            if (chunk_slot[c - 1] >= 0) read_graph_events(chunk_slot[c - 1]); // chunk c (the other executable) is running, c + 1 not yet launched
        }
        ++c;
    }
    CK(cudaEventRecord(h->ev_stop, h->stream));
    CK(cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "the CG loop faulted on the device: %s", cudaGetErrorString(se));
    const DevState &s = h->h_st[2];
    rc = check_device_error(h, s);
    if (rc != LAMCG_OK) return rc;

    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
    double gemv_ms = 0.0;
    int gemv_timed = 0; // GEMV launches the sum covers
    if (graph_timed) {
        if (c >= 1 && chunk_slot[c - 1] >= 0) read_graph_events(chunk_slot[c - 1]); // the last chunk (the stream is idle now)
        const int cnt = std::min((int)graph_gemv_ms.size(), s.iters_done - first_iter); // launches after `done` are no-ops: not counted
        for (int i = 0; i < cnt; ++i)
            if (graph_gemv_ms[i] >= 0.f) { gemv_ms += graph_gemv_ms[i]; ++gemv_timed; }
    } else if (time_gemv) {
        const int cnt = std::min(timed_iters, s.iters_done - first_iter);
        gemv_timed = cnt;
        for (int i = 0; i < cnt; ++i) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, h->gemv_events[2 * i], h->gemv_events[2 * i + 1]));
            gemv_ms += t;
        }
    }
    h->last_hist_count = h->opt_history ? std::min(s.iters_done, h->hist_cap) : 0;
    // x, r, the p of the last iteration, beta and rr are still on the device: a solve that ran out of iterations can go on
    h->resumable = !s.converged && !s.breakdown && s.iters_done >= 1 && s.iters_done == max_total;
    h->done_iters = s.iters_done;
    if (out) {
        out->converged = s.converged;
        out->iterations = s.converged ? s.iters_done : (max_total < 0 ? 1 : max_total + 1);
        out->rel_residual = std::sqrt(s.rr_final / s.bb);
        out->solve_seconds = ms * 1e-3;
        out->gemv_seconds = gemv_ms * 1e-3;
        out->iterations_run = s.iters_done;
        out->kernel_launches = launches;
        out->numerical_breakdown = s.breakdown;
        out->gemv_launches_timed = gemv_timed;
    }
    return LAMCG_OK;
}

} // namespace

// ---- checkpoint / restart of a long solve ------------------------------------
// One file per rank: a fixed header, then this rank's slices of x, r and of the p of the last executed
// iteration, then the residual history kept so far.  Everything else the loop needs (A, b) is the system itself.
namespace {

struct CkptHeader {
    char magic[8];            // "LAMCGCK1"
    uint64_t n, local_rows, row_offset;
    int32_t dtype, rank, nranks, iters_done;
    double bb, rr, alpha_last, beta_last, eps;
    int32_t hist_count, reserved;
};
const char kCkptMagic[8] = {'L', 'A', 'M', 'C', 'G', 'C', 'K', '1'};

int write_full(int fd, const void *buf, size_t bytes)
{
    const char *c = static_cast<const char *>(buf);
    while (bytes > 0) {
        ssize_t put = write(fd, c, bytes);
        if (put < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        c += put;
        bytes -= (size_t)put;
    }
    return 0;
}
} // namespace


// =================================================================================================
// extern "C" surface
// =================================================================================================
extern "C" {

const char *lamcg_version(void) { return "lamcg-b200 0.1 (sm_100a)"; }

const char *lamcg_last_error(const lamcg_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lamcg_create_ranked(lamcg_t **out, int device, int rank, int nranks)
{
    if (!out) return LAMCG_ERR_INVALID;
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        g_create_error = "invalid rank/nranks";
        return LAMCG_ERR_INVALID;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        cudaGetLastError();
        return LAMCG_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device ordinal out of range";
        return LAMCG_ERR_INVALID;
    }
    lamcg *h = new lamcg();
    h->device = device;
    h->rank = rank;
    h->nranks = nranks;
    auto bail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        cudaGetLastError();
        lamcg_destroy(h); // releases whatever was created so far
        return (int)LAMCG_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        g_create_error = "this library is built for sm_100a (Blackwell B200) only";
        delete h;
        return LAMCG_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    h->clock_khz = prop.clockRate > 0 ? prop.clockRate : 1965000;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaMalloc(&h->st, sizeof(DevState))) != cudaSuccess) return bail("cudaMalloc(state)", e);
    if ((e = cudaMemset(h->st, 0, sizeof(DevState))) != cudaSuccess) return bail("cudaMemset(state)", e);
    if ((e = cudaMalloc(&h->partials, 2 * kMaxGrid * sizeof(double))) != cudaSuccess) return bail("cudaMalloc(partials)", e);
    if ((e = cudaMallocHost(&h->h_st, 3 * sizeof(DevState))) != cudaSuccess) return bail("cudaMallocHost(status)", e);
    cudaEventCreate(&h->ev_start);
    cudaEventCreate(&h->ev_stop);
    cudaEventCreateWithFlags(&h->ev_ring[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_ring[1], cudaEventDisableTiming);
    h->opt_gemv_variant = env_ll("gemv_variant", 0);
    h->opt_loop_mode = env_ll("loop_mode", 0);
    h->opt_chunk_iters = env_ll("chunk_iters", 16);
    h->opt_time_gemv = env_ll("time_gemv", 0);
    h->opt_history = env_ll("history", 1);
    h->opt_gemv_ctas_per_sm = env_ll("gemv_ctas_per_sm", 0);
    h->opt_persist_rows_smem = env_ll("persist_rows_smem", -1);
    h->opt_persist_variant = env_ll("persist_variant", 0);
    h->opt_persist_poll_delay = env_ll("persist_poll_delay", 650);
    h->opt_fuse_updates = env_ll("fuse_updates", 1);
    h->opt_ingest_threads = env_ll("ingest_threads", 8);
    h->opt_ingest_chunk_bytes = env_ll("ingest_chunk_bytes", 4ll << 20);
    h->opt_peer_timeout_s = env_ll("peer_timeout_s", 600);
    h->opt_persist_grid = env_ll("persist_grid", 0);
    h->opt_matrix_f32 = env_ll("matrix_f32", 0) == 1 ? 1 : 0;
    h->asz = h->opt_matrix_f32 ? sizeof(float) : sizeof(double);
    *out = h;
    return LAMCG_OK;
}

int lamcg_create(lamcg_t **out, int device) { return lamcg_create_ranked(out, device, 0, 1); }

int lamcg_create_typed(lamcg_t **out, int device, int rank, int nranks, int dtype)
{
    if (dtype != 0 && dtype != 1) {
        g_create_error = "dtype must be 0 (fp64) or 1 (fp32)";
        if (out) *out = nullptr;
        return LAMCG_ERR_INVALID;
    }
    int rc = lamcg_create_ranked(out, device, rank, nranks);
    if (rc != LAMCG_OK) return rc;
    (*out)->dtype = dtype;
    (*out)->esz = dtype == 0 ? sizeof(double) : sizeof(float);
    if (dtype != 0) (*out)->opt_matrix_f32 = 0; // an fp32 handle already stores fp32
    (*out)->asz = (*out)->opt_matrix_f32 ? sizeof(float) : (*out)->esz;
    return LAMCG_OK;
}

void lamcg_destroy(lamcg_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    // A CUDA graph that captured NCCL kernels keeps the communicator referenced: ncclCommDestroy
    // blocks until such graphs are gone, so the executable graph goes first.
    destroy_graphs(h);
    if (h->nccl) nccl_api().CommDestroy(h->nccl);
    close_peer_handles(h);
    cudaFree(h->peer_base);
    cudaFree(h->persist_ll);
    cudaFree(h->narrow_stats);
    if (h->ingest_pool) cudaFreeHost(h->ingest_pool);
    free_system(h);
    for (cudaEvent_t e : h->gemv_events) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k)
        for (cudaEvent_t e : h->graph_events[k]) cudaEventDestroy(e);
    cudaFree(h->hist);
    cudaFree(h->st);
    cudaFree(h->partials);
    cudaFreeHost(h->h_st);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    if (h->ev_ring[0]) cudaEventDestroy(h->ev_ring[0]);
    if (h->ev_ring[1]) cudaEventDestroy(h->ev_ring[1]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int lamcg_set_option(lamcg_t *h, const char *key, long long value)
{
    if (!h || !key) return LAMCG_ERR_INVALID;
    std::string k(key);
    const long long old_variant = h->opt_gemv_variant;
    if (k == "gemv_variant") h->opt_gemv_variant = value;
    else if (k == "loop_mode") h->opt_loop_mode = value;
    else if (k == "chunk_iters") h->opt_chunk_iters = value;
    else if (k == "time_gemv") h->opt_time_gemv = value;
    else if (k == "history") h->opt_history = value;
    else if (k == "gemv_ctas_per_sm") h->opt_gemv_ctas_per_sm = value;
    else if (k == "persist_rows_smem") h->opt_persist_rows_smem = value;
    else if (k == "persist_variant") h->opt_persist_variant = value;
    else if (k == "persist_l2_keep_mb") h->opt_persist_l2_keep_mb = value;
    else if (k == "persist_poll_delay") h->opt_persist_poll_delay = value;
    else if (k == "ingest_threads") h->opt_ingest_threads = value;
    else if (k == "ingest_chunk_bytes") h->opt_ingest_chunk_bytes = value;
    else if (k == "peer_timeout_s") { h->opt_peer_timeout_s = value; h->pv.timeout_cycles = peer_timeout_cycles(h); }
    else if (k == "persist_grid") h->opt_persist_grid = value;
    else if (k == "debug_persist_fail") h->opt_debug_persist_fail = value;
    else if (k == "spd_simt") h->opt_spd_simt = value;
    else if (k == "fuse_updates") h->opt_fuse_updates = value;
    else if (k == "loop_profile") h->opt_loop_profile = value;
    else if (k == "matrix_f32") {
        if (value != 0 && value != 1) return h->fail(LAMCG_ERR_INVALID, "matrix_f32 is 0 or 1");
        if (value && h->dtype != 0) return h->fail(LAMCG_ERR_INVALID, "matrix_f32 applies to fp64 handles (an fp32 handle already stores fp32)");
        if (value != h->opt_matrix_f32) {
            // the matrix block changes its element size: whatever system is loaded is dropped and has to be loaded again
            free_system(h);
            h->resumable = false;
            h->opt_matrix_f32 = value;
            h->asz = value ? sizeof(float) : h->esz;
            h->last_inexact = h->last_overflow = 0;
        }
    }
    else return h->fail(LAMCG_ERR_INVALID, "unknown option '%s'", key);
    destroy_graphs(h);
    if (h->alloc_n) {
        CK(cudaSetDevice(h->device));
        const int rc = make_plan(h);
        if (rc != LAMCG_OK && k == "gemv_variant") h->opt_gemv_variant = old_variant; // refused: the loaded system keeps its working plan
        return rc;
    }
    return LAMCG_OK;
}

int lamcg_get_info(const lamcg_t *h, lamcg_info *out)
{
    if (!h || !out) return LAMCG_ERR_INVALID;
    out->n = h->n;
    out->local_rows = h->local_rows;
    out->row_offset = h->row_offset;
    out->lda = h->lda;
    out->rank = h->rank;
    out->nranks = h->nranks;
    out->device = h->device;
    out->sm_count = h->sm_count;
    out->comm_mode = h->comm_mode;
    out->has_matrix = h->has_matrix;
    out->has_rhs = h->has_rhs;
    out->gemv_variant = h->plan.variant;
    out->gemv_grid = h->plan.grid;
    out->gemv_block = h->plan.block;
    out->gemv_smem_bytes = (int)h->plan.smem;
    out->dtype = h->dtype;
    out->ingest_threads = h->last_ingest_threads;
    out->ingest_chunks = (int)std::min<long long>(h->last_ingest_chunks, INT_MAX);
    out->matrix_elem_bytes = (int)h->asz;
    out->matrix_f32_inexact = h->last_inexact;
    out->matrix_f32_overflow = h->last_overflow;
    return LAMCG_OK;
}

// ---- comm -----------------------------------------------------------------------------------
int lamcg_comm_nccl_unique_id(void *id_out)
{
    if (!id_out) return LAMCG_ERR_INVALID;
    const char *why = "";
    if (!nccl_api().load(&why)) {
        g_create_error = why;
        return LAMCG_ERR_COMM;
    }
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == LAMCG_NCCL_ID_BYTES, "ncclUniqueId size");
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) {
        g_create_error = "ncclGetUniqueId failed";
        return LAMCG_ERR_COMM;
    }
    memcpy(id_out, &id, sizeof id);
    return LAMCG_OK;
}

int lamcg_comm_init_nccl(lamcg_t *h, const void *id)
{
    if (!h || !id) return LAMCG_ERR_INVALID;
    if (h->nranks == 1) return LAMCG_OK;
    const char *why = "";
    if (!nccl_api().load(&why)) return h->fail(LAMCG_ERR_COMM, "%s", why);
    CK(cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    NCK(nccl_api().CommInitRank(&h->nccl, h->nranks, uid, h->rank));
    h->comm_mode = kCommNccl;
    return LAMCG_OK;
}

namespace {
struct PeerHandleWire { // what travels between ranks: LAMCG_PEER_HANDLE_BYTES
    cudaIpcMemHandle_t ipc;
    unsigned long long bytes, n;
    int rank, pid;
};
static_assert(sizeof(PeerHandleWire) <= LAMCG_PEER_HANDLE_BYTES, "peer handle too large");
} // namespace

int lamcg_comm_peer_export(lamcg_t *h, size_t n, void *handle_out)
{
    if (!h || !handle_out || n == 0) return LAMCG_ERR_INVALID;
    if (h->nranks > kMaxRanks) return h->fail(LAMCG_ERR_INVALID, "peer exchange supports at most %d ranks", kMaxRanks);
    if (h->comm_mode != kCommNone) return h->fail(LAMCG_ERR_STATE, "a communicator is already initialised");
    CK(cudaSetDevice(h->device));
    if (h->peer_base) { cudaFree(h->peer_base); h->peer_base = nullptr; }
    const size_t lda = (n + 15) / 16 * 16;
    const size_t hdr = (sizeof(PeerHeader) + 255) / 256 * 256;
    h->pv = PeerView{};
    h->pv.timeout_cycles = peer_timeout_cycles(h);
    h->pv.off_p[0] = (long long)hdr;
    h->pv.off_p[1] = (long long)(hdr + lda * 8);   // sized for fp64; fp32 handles use the first half of each buffer
    h->pv.off_xg[0] = (long long)(hdr + 2 * lda * 8);
    h->pv.off_xg[1] = (long long)(hdr + 3 * lda * 8);
    h->peer_bytes = hdr + 4 * lda * 8;
    h->peer_n = n;
    CK(cudaMalloc(&h->peer_base, h->peer_bytes));
    CK(cudaMemset(h->peer_base, 0, h->peer_bytes));
    CK(cudaDeviceSynchronize());
    PeerHandleWire w{};
    CK(cudaIpcGetMemHandle(&w.ipc, h->peer_base));
    w.bytes = h->peer_bytes;
    w.n = n;
    w.rank = h->rank;
    w.pid = (int)getpid();
    memset(handle_out, 0, LAMCG_PEER_HANDLE_BYTES);
    memcpy(handle_out, &w, sizeof w);
    return LAMCG_OK;
}

int lamcg_comm_init_peer(lamcg_t *h, const void *all_handles)
{
    if (!h || !all_handles) return LAMCG_ERR_INVALID;
    if (h->nranks == 1) return LAMCG_OK;
    if (!h->peer_base) return h->fail(LAMCG_ERR_STATE, "lamcg_comm_peer_export must be called first");
    CK(cudaSetDevice(h->device));
    const unsigned char *blob = static_cast<const unsigned char *>(all_handles);
    for (int r = 0; r < h->nranks; ++r) {
        PeerHandleWire w;
        memcpy(&w, blob + (size_t)r * LAMCG_PEER_HANDLE_BYTES, sizeof w);
        if (w.rank != r || w.n != h->peer_n || w.bytes != h->peer_bytes) {
            close_peer_handles(h);
            return h->fail(LAMCG_ERR_COMM, "peer handle %d is inconsistent (rank %d, n %llu, bytes %llu)", r, w.rank, w.n, w.bytes);
        }
        if (r == h->rank) {
            h->pv.base[r] = h->peer_base;
        } else {
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, w.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                close_peer_handles(h); // comm_mode stays kCommNone, so destroy would not close what was opened so far
                return h->fail(LAMCG_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s (peer access over NVLink is required)", r,
                               cudaGetErrorString(e));
            }
            h->pv.base[r] = static_cast<unsigned char *>(ptr);
        }
    }
    h->pv.me = h->rank;
    h->pv.nranks = h->nranks;
    h->comm_mode = kCommPeer;
    destroy_graphs(h);
    return LAMCG_OK;
}

// ---- system ---------------------------------------------------------------------------------
int lamcg_generate_matrix(lamcg_t *h, size_t rows, size_t cols)
{
    if (!h) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (rows != cols) return h->fail(LAMCG_ERR_SHAPE, "Matrix has to be square");
    int rc = alloc_system(h, rows);
    if (rc != LAMCG_OK) return rc;
    if (h->local_rows > 0) {
        const long long total = (long long)h->local_rows * (long long)(h->lda / 2);
        const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
        if (h->dtype == 0 && !mixed_storage(h))
            generate_matrix_kernel<double><<<grid, 256, 0, h->stream>>>(h->A, (long long)h->local_rows, (long long)h->n, (long long)h->lda,
                                                                       (long long)h->row_offset);
        else
            generate_matrix_kernel<float><<<grid, 256, 0, h->stream>>>(h->A, (long long)h->local_rows, (long long)h->n, (long long)h->lda,
                                                                      (long long)h->row_offset);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    h->last_inexact = h->last_overflow = 0; // 0, 1 and 2 are fp32 numbers
    h->has_matrix = true;
    h->has_rhs = false;
    return LAMCG_OK;
}

int lamcg_generate_rhs(lamcg_t *h)
{
    if (!h) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "generate_rhs before a matrix exists");
    CK(cudaSetDevice(h->device));
    const int grid = (int)std::min<size_t>((h->lda + 255) / 256, (size_t)h->sm_count * 4);
    if (h->dtype == 0) fill_kernel<double><<<grid, 256, 0, h->stream>>>(h->b_full, (long long)h->n, (long long)h->lda, 1.0);
    else fill_kernel<float><<<grid, 256, 0, h->stream>>>(h->b_full, (long long)h->n, (long long)h->lda, 1.0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->has_rhs = true;
    return LAMCG_OK;
}

int lamcg_set_matrix(lamcg_t *h, const void *A, size_t n, int layout)
{
    if (!h || !A) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    int rc = alloc_system(h, n);
    if (rc != LAMCG_OK) return rc;
    const char *src = layout == 0 ? static_cast<const char *>(A) + h->row_offset * n * h->esz : static_cast<const char *>(A);
    if (h->lda != n) CK(cudaMemsetAsync(h->A, 0, h->local_rows * h->lda * h->asz, h->stream));
    if (mixed_storage(h)) {
        // fp64 source (host or device) -> fp32 block: row chunks of at most 64 MB through a device staging buffer
        rc = narrow_begin(h);
        if (rc != LAMCG_OK) return rc;
        const size_t chunk_rows = std::min(std::max<size_t>(1, ((size_t)64 << 20) / (n * sizeof(double))), std::max<size_t>(h->local_rows, 1));
        double *stage = nullptr;
        CK(cudaMalloc(&stage, chunk_rows * n * sizeof(double)));
        cudaError_t e = cudaSuccess;
        for (size_t r = 0; r < h->local_rows && e == cudaSuccess; r += chunk_rows) {
            const size_t nr = std::min(chunk_rows, h->local_rows - r);
            e = cudaMemcpyAsync(stage, src + r * n * sizeof(double), nr * n * sizeof(double), cudaMemcpyDefault, h->stream);
            if (e == cudaSuccess) e = narrow_rows(h, stage, n, r, nr, h->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(stage);
        if (e != cudaSuccess) return h->fail(LAMCG_ERR_CUDA, "narrowing the matrix to fp32 failed: %s", cudaGetErrorString(e));
        rc = narrow_end(h);
        if (rc != LAMCG_OK) return rc;
    } else if (h->local_rows > 0)
        CK(cudaMemcpy2DAsync(h->A, h->lda * h->esz, src, n * h->esz, n * h->esz, h->local_rows, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->has_matrix = true;
    h->has_rhs = false;
    return LAMCG_OK;
}

int lamcg_set_rhs(lamcg_t *h, const void *b, size_t n)
{
    if (!h || !b) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "set_rhs before a matrix exists");
    if (n != h->n) return h->fail(LAMCG_ERR_SHAPE, "Size of right hand side does not match the matrix");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->b_full, b, n * h->esz, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->has_rhs = true;
    return LAMCG_OK;
}

int lamcg_load_matrix(lamcg_t *h, const char *path)
{
    if (!h || !path) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    int fd = open(path, O_RDONLY);
    if (fd < 0) return h->fail(LAMCG_ERR_IO, "Cannot open %s: %s", path, strerror(errno));
    size_t rows = 0, cols = 0;
    int rc = read_header(h, fd, path, &rows, &cols);
    if (rc != LAMCG_OK) { close(fd); return rc; }
    if (rows != cols) { close(fd); return h->fail(LAMCG_ERR_SHAPE, "Matrix has to be square"); }
    struct stat sb;
    if (fstat(fd, &sb) == 0 && (unsigned long long)sb.st_size < 16ull + (unsigned long long)rows * cols * (unsigned long long)h->esz) {
        close(fd);
        return h->fail(LAMCG_ERR_IO, "%s: file is shorter than its %zu x %zu header promises", path, rows, cols);
    }
    rc = alloc_system(h, rows);
    if (rc != LAMCG_OK) { close(fd); return rc; }
    const size_t n = h->n;
    // Chunked, multi-threaded ingest: T reader threads pull row chunks off a shared counter; each owns its own stream and
    // kIngestSlots pinned staging buffers out of a pool that lives in the handle (pinning memory costs ~0.3 s per GB: round 1
    // allocated 2 x 32 MB per thread on EVERY load, which is why 8 threads lost to 4 on a 3 GB file).  pread (page cache / disk
    // -> pinned) of one chunk overlaps the async 2-D H2D copies of the thread's previous chunks.  One thread tops out near
    // 6 GB/s (a single core's copy rate out of the page cache, the rate the reference's fread reaches into pageable memory);
    // T threads scale that until PCIe is the limit.  All offsets are 64-bit (the reference's MPI-IO count is an int: n = 50000
    // on one rank reads garbage, TESTS/BEST_RESULTS:114).
    const size_t row_bytes = n * h->esz; // file elements have the handle's type, like the reference's sizeof(FloatingType)
    const size_t chunk_bytes = (size_t)std::max<long long>(row_bytes, std::min<long long>(h->opt_ingest_chunk_bytes, 256ll << 20));
    size_t chunk_rows = std::max<size_t>(1, chunk_bytes / row_bytes);
    chunk_rows = std::min(chunk_rows, std::max<size_t>(h->local_rows, 1));
    const size_t nchunks = (h->local_rows + chunk_rows - 1) / chunk_rows;
    int T = (int)std::max<long long>(1, std::min<long long>(h->opt_ingest_threads, kIngestMaxThreads));
    T = (int)std::min<size_t>((size_t)T, std::max<size_t>(nchunks, 1));
    const size_t slot_bytes = chunk_rows * row_bytes;
    rc = ensure_ingest_pool(h, (size_t)T * kIngestSlots * slot_bytes);
    if (rc != LAMCG_OK) { close(fd); return rc; }
    if (h->lda != n) CK(cudaMemsetAsync(h->A, 0, h->local_rows * h->lda * h->asz, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // option matrix_f32: a chunk lands in a device staging slot (one per pinned slot) and is narrowed into the block by a kernel
    // on the reader's stream; the slot's event then covers both buffers
    const bool mixed = mixed_storage(h);
    char *dev_stage = nullptr;
    if (mixed) {
        rc = narrow_begin(h);
        if (rc != LAMCG_OK) { close(fd); return rc; }
        cudaError_t e = cudaMalloc(&dev_stage, (size_t)T * kIngestSlots * slot_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            close(fd);
            return h->fail(LAMCG_ERR_NOMEM, "cudaMalloc of the ingest staging buffers failed: %s", cudaGetErrorString(e));
        }
    }
    std::atomic<int> status{LAMCG_OK};
    std::atomic<size_t> next_chunk{0};
    std::string first_error;
    std::mutex err_mu;
    auto worker = [&](int t) {
        auto failw = [&](int code, const std::string &msg) {
            std::lock_guard<std::mutex> lk(err_mu);
            if (status.load() == LAMCG_OK) { status.store(code); first_error = msg; }
        };
        if (cudaSetDevice(h->device) != cudaSuccess) return failw(LAMCG_ERR_CUDA, "cudaSetDevice failed in ingest thread");
        cudaStream_t st = nullptr;
        cudaEvent_t done[kIngestSlots] = {};
        bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < kIngestSlots && ok; ++i) ok = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) failw(LAMCG_ERR_CUDA, "stream / event creation failed in ingest thread");
        char *stage0 = h->ingest_pool + (size_t)t * kIngestSlots * slot_bytes;
        int slot = 0;
        while (ok && status.load() == LAMCG_OK) {
            const size_t c = next_chunk.fetch_add(1);
            if (c >= nchunks) break;
            const size_t r = c * chunk_rows;
            const size_t nr = std::min(chunk_rows, h->local_rows - r);
            char *stage = stage0 + (size_t)slot * slot_bytes;
            cudaEventSynchronize(done[slot]); // the H2D copy that last used this slot has drained
            const off_t off = (off_t)16 + (off_t)((h->row_offset + r) * row_bytes);
            if (pread_full(fd, stage, nr * row_bytes, off) != 0) {
                failw(LAMCG_ERR_IO, std::string(path) + ": short read in rows " + std::to_string(h->row_offset + r) + ".." +
                                        std::to_string(h->row_offset + r + nr));
                break;
            }
            cudaError_t e;
            if (mixed) {
                char *dstage = dev_stage + ((size_t)t * kIngestSlots + slot) * slot_bytes;
                e = cudaMemcpyAsync(dstage, stage, nr * row_bytes, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = narrow_rows(h, reinterpret_cast<const double *>(dstage), n, r, nr, st);
            } else
                e = cudaMemcpy2DAsync(h->A + r * h->lda * h->esz, h->lda * h->esz, stage, row_bytes, row_bytes, nr, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { failw(LAMCG_ERR_CUDA, std::string("matrix upload failed: ") + cudaGetErrorString(e)); break; }
            cudaEventRecord(done[slot], st);
            slot = (slot + 1) % kIngestSlots;
        }
        if (st) {
            if (cudaStreamSynchronize(st) != cudaSuccess) failw(LAMCG_ERR_CUDA, "matrix upload failed");
            cudaStreamDestroy(st);
        }
        for (int i = 0; i < kIngestSlots; ++i)
            if (done[i]) cudaEventDestroy(done[i]);
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(worker, t);
        worker(0);
        for (auto &th : pool) th.join();
    }
    h->last_ingest_threads = T;
    h->last_ingest_chunks = (long long)nchunks;
    close(fd);
    CK(cudaSetDevice(h->device));
    if (dev_stage) cudaFree(dev_stage);
    if (status.load() != LAMCG_OK) return h->fail(status.load(), "%s", first_error.c_str());
    if (mixed) {
        rc = narrow_end(h);
        if (rc != LAMCG_OK) return rc;
    }
    h->has_matrix = true;
    h->has_rhs = false;
    return LAMCG_OK;
}

int lamcg_load_rhs(lamcg_t *h, const char *path)
{
    if (!h || !path) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "load_rhs before a matrix exists");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return h->fail(LAMCG_ERR_IO, "Cannot open %s: %s", path, strerror(errno));
    size_t rows = 0, cols = 0;
    int rc = read_header(h, fd, path, &rows, &cols);
    if (rc != LAMCG_OK) { close(fd); return rc; }
    if (cols != 1) { close(fd); return h->fail(LAMCG_ERR_SHAPE, "The file does not contain a valid rhs"); }
    if (rows != h->n) { close(fd); return h->fail(LAMCG_ERR_SHAPE, "Size of right hand side does not match the matrix"); }
    std::vector<char> b(rows * h->esz);
    if (pread_full(fd, b.data(), rows * h->esz, 16) != 0) {
        close(fd);
        return h->fail(LAMCG_ERR_IO, "%s: short read", path);
    }
    close(fd);
    return lamcg_set_rhs(h, b.data(), rows);
}

// ---- solve ----------------------------------------------------------------------------------
int lamcg_solve(lamcg_t *h, int max_iters, double rel_error, lamcg_result *out)
{
    if (!h) return LAMCG_ERR_INVALID;
    if (!h->has_matrix || !h->has_rhs) return h->fail(LAMCG_ERR_STATE, "solve needs a matrix and a right hand side");
    if (h->nranks > 1 && h->comm_mode == kCommNone) return h->fail(LAMCG_ERR_STATE, "multi-rank solve before lamcg_comm_init_*");
    CK(cudaSetDevice(h->device));
    h->resumable = false;
    int rc = ensure_hist(h, max_iters, 0);
    if (rc != LAMCG_OK) return rc;

    const int loop_mode = resolve_loop_mode(h);
    int loop_mode_used = loop_mode;
    if (loop_mode == kLoopPersistent) {
        rc = solve_persistent(h, max_iters, rel_error, out);
        if (rc != kPersistUnavailable) return rc;
        if ((int)h->opt_loop_mode == kLoopPersistent) // asked for explicitly: report (message already set)
            return h->err.find("rows per CTA") != std::string::npos ? LAMCG_ERR_INVALID : LAMCG_ERR_CUDA;
        loop_mode_used = kLoopGraph; // chosen by size: the graph loop does the same job
        h->err.clear();              // ... and the solve is not a failure
    }

    InitArgs ia;
    ia.st = h->st;
    ia.b_full = h->b_full;
    ia.x = h->x;
    ia.r = h->r;
    ia.Ap = h->Ap;
    ia.p_full = p_ptr(h, 0);
    ia.seq_base = h->cur_seq_base = h->seq_next;
    h->seq_next += (unsigned long long)std::max(max_iters, 0) + 2ull;
    ia.n = (long long)h->n;
    ia.lda = (long long)h->lda;
    ia.rows = (long long)h->local_rows;
    ia.row_offset = (long long)h->row_offset;
    ia.eps = rel_error;
    ia.max_iters = max_iters;
    ia.hist_cap = h->opt_history ? h->hist_cap : 0;
    if (h->dtype == 0) init_solve_kernel<double><<<1, 1024, 0, h->stream>>>(ia);
    else init_solve_kernel<float><<<1, 1024, 0, h->stream>>>(ia);
    CK(cudaGetLastError());
    return run_loop(h, loop_mode_used, 0, max_iters, out, 1);
}

int lamcg_solve_resume(lamcg_t *h, int more_iters, double rel_error, lamcg_result *out)
{
    if (!h) return LAMCG_ERR_INVALID;
    if (more_iters < 1) return h->fail(LAMCG_ERR_INVALID, "lamcg_solve_resume: more_iters must be >= 1");
    if (!h->resumable)
        return h->fail(LAMCG_ERR_STATE, "nothing to resume: the last solve converged, broke down, ran in the persistent loop "
                                        "(set loop_mode 2) or the system changed since; or no checkpoint was loaded");
    if (h->done_iters > INT_MAX - more_iters) return h->fail(LAMCG_ERR_INVALID, "iteration count overflows int");
    CK(cudaSetDevice(h->device));
    const int k = h->done_iters, total = k + more_iters;
    int rc = ensure_hist(h, total, std::min(k, h->hist_cap));
    if (rc != LAMCG_OK) return rc;
    int loop_mode = resolve_loop_mode(h);
    if (loop_mode == kLoopPersistent) loop_mode = kLoopGraph; // the one-kernel loop always starts from x = 0
    h->seq_next = std::max(h->seq_next, h->cur_seq_base + (unsigned long long)total + 2ull);
    h->resumable = false;
    // the deferred p update of iteration k (0-based k-1), then carry on with iteration index k
    VecArgs v = vec_args(h, (k - 1) & 1);
    const int vg = vec_grid(h);
    const int hist_cap = h->opt_history ? h->hist_cap : 0;
    if (h->dtype == 0) resume_kernel<double><<<vg, kVecThreads, 0, h->stream>>>(v, total, rel_error, hist_cap);
    else resume_kernel<float><<<vg, kVecThreads, 0, h->stream>>>(v, total, rel_error, hist_cap);
    CK(cudaGetLastError());
    if (h->comm_mode == kCommNccl) {
        rc = allgather_vec(h, h->p_full);
        if (rc != LAMCG_OK) return rc;
    }
    return run_loop(h, loop_mode, k, total, out, 1);
}

int lamcg_checkpoint_save(lamcg_t *h, const char *path)
{
    if (!h || !path) return LAMCG_ERR_INVALID;
    if (!h->resumable) return h->fail(LAMCG_ERR_STATE, "no resumable state to checkpoint (the last solve must have stopped on max_iters in the stream or graph loop)");
    CK(cudaSetDevice(h->device));
    const int k = h->done_iters;
    CK(cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    const size_t slice = h->local_rows * h->esz;
    const int hist_count = h->opt_history ? std::min(k, h->hist_cap) : 0;
    std::vector<char> buf(3 * slice + (size_t)hist_count * sizeof(double));
    if (slice) {
        CK(cudaMemcpyAsync(buf.data(), h->x, slice, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(buf.data() + slice, h->r, slice, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(buf.data() + 2 * slice, p_ptr(h, (k - 1) & 1) + h->row_offset * h->esz, slice, cudaMemcpyDeviceToHost, h->stream));
    }
    if (hist_count) CK(cudaMemcpyAsync(buf.data() + 3 * slice, h->hist, (size_t)hist_count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const DevState &s = h->h_st[2];
    CkptHeader hd{};
    memcpy(hd.magic, kCkptMagic, 8);
    hd.n = h->n; hd.local_rows = h->local_rows; hd.row_offset = h->row_offset;
    hd.dtype = h->dtype; hd.rank = h->rank; hd.nranks = h->nranks; hd.iters_done = k;
    hd.bb = s.bb; hd.rr = s.rr_final; hd.alpha_last = s.alpha_last; hd.beta_last = s.beta_last; hd.eps = s.eps;
    hd.hist_count = hist_count;
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return h->fail(LAMCG_ERR_IO, "%s: cannot open for writing: %s", path, strerror(errno));
    const bool ok = write_full(fd, &hd, sizeof hd) == 0 && write_full(fd, buf.data(), buf.size()) == 0;
    const bool closed = close(fd) == 0;
    if (!ok || !closed) return h->fail(LAMCG_ERR_IO, "%s: short write", path);
    return LAMCG_OK;
}

int lamcg_checkpoint_load(lamcg_t *h, const char *path)
{
    if (!h || !path) return LAMCG_ERR_INVALID;
    if (!h->has_matrix || !h->has_rhs) return h->fail(LAMCG_ERR_STATE, "load the system (matrix and rhs) before its checkpoint");
    if (h->nranks > 1 && h->comm_mode == kCommNone) return h->fail(LAMCG_ERR_STATE, "multi-rank checkpoint load before lamcg_comm_init_*");
    h->resumable = false;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return h->fail(LAMCG_ERR_IO, "%s: cannot open: %s", path, strerror(errno));
    CkptHeader hd;
    if (pread_full(fd, &hd, sizeof hd, 0) != 0 || memcmp(hd.magic, kCkptMagic, 8) != 0) {
        close(fd);
        return h->fail(LAMCG_ERR_IO, "%s: not a lamcg checkpoint", path);
    }
    if (hd.n != h->n || hd.local_rows != h->local_rows || hd.row_offset != h->row_offset || hd.dtype != h->dtype || hd.rank != h->rank ||
        hd.nranks != h->nranks || hd.iters_done < 1 || hd.hist_count < 0 || hd.hist_count > hd.iters_done) {
        close(fd);
        return h->fail(LAMCG_ERR_SHAPE, "%s: checkpoint of rank %d/%d, n = %llu, dtype %d does not match this handle (rank %d/%d, n = %zu, dtype %d)", path,
                       hd.rank, hd.nranks, (unsigned long long)hd.n, hd.dtype, h->rank, h->nranks, h->n, h->dtype);
    }
    const size_t slice = h->local_rows * h->esz;
    std::vector<char> buf(3 * slice + (size_t)hd.hist_count * sizeof(double));
    const int rd = buf.empty() ? 0 : pread_full(fd, buf.data(), buf.size(), (off_t)sizeof hd);
    close(fd);
    if (rd != 0) return h->fail(LAMCG_ERR_IO, "%s: truncated checkpoint", path);
    CK(cudaSetDevice(h->device));
    const int k = hd.iters_done;
    int rc = ensure_hist(h, k, 0);
    if (rc != LAMCG_OK) return rc;
    if (slice) {
        CK(cudaMemcpyAsync(h->x, buf.data(), slice, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->r, buf.data() + slice, slice, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(p_ptr(h, (k - 1) & 1) + h->row_offset * h->esz, buf.data() + 2 * slice, slice, cudaMemcpyHostToDevice, h->stream));
    }
    const int hist_keep = h->opt_history ? std::min(hd.hist_count, h->hist_cap) : 0;
    if (hist_keep) CK(cudaMemcpyAsync(h->hist, buf.data() + 3 * slice, (size_t)hist_keep * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // the scalar state exactly as K3 of iteration k left it (both parity slots hold the latest values)
    DevState s{};
    s.bb = hd.bb;
    s.rr[0] = s.rr[1] = s.rr_final = hd.rr;
    s.alpha_last = hd.alpha_last;
    s.beta_last = hd.beta_last;
    s.eps = hd.eps;
    s.iter[0] = s.iter[1] = s.iters_done = k;
    s.max_iters = k;
    s.done = 1;
    s.hist_cap = h->opt_history ? h->hist_cap : 0;
    s.seq_base = h->cur_seq_base = h->seq_next;
    h->seq_next += (unsigned long long)k + 2ull;
    h->h_st[2] = s;
    CK(cudaMemcpyAsync(h->st, &h->h_st[2], sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->last_hist_count = hist_keep;
    h->done_iters = k;
    h->resumable = true;
    return LAMCG_OK;
}

int lamcg_get_residual_history(lamcg_t *h, double *out, int capacity)
{
    if (!h || !out || capacity < 0) return LAMCG_ERR_INVALID;
    const int cnt = std::min(capacity, h->last_hist_count);
    if (cnt > 0) {
        CK(cudaSetDevice(h->device));
        CK(cudaMemcpyAsync(out, h->hist, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return cnt;
}

int lamcg_get_solution_local(lamcg_t *h, void *x_local)
{
    if (!h || !x_local) return LAMCG_ERR_INVALID;
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no system");
    CK(cudaSetDevice(h->device));
    if (h->local_rows) CK(cudaMemcpyAsync(x_local, h->x, h->local_rows * h->esz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LAMCG_OK;
}

int lamcg_get_solution(lamcg_t *h, void *x)
{
    if (!h || !x) return LAMCG_ERR_INVALID;
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no system");
    if (h->nranks == 1) return lamcg_get_solution_local(h, x);
    CK(cudaSetDevice(h->device));
    if (h->comm_mode == kCommPeer) {
        const unsigned long long gseq = ++h->gather_seq;
        const int buf = (int)(gseq & 1ull);
        if (h->dtype == 0)
            peer_gather_put_kernel<double><<<vec_grid(h), kVecThreads, 0, h->stream>>>(h->pv, h->x, (long long)h->local_rows,
                                                                                      (long long)h->row_offset, buf, gseq, h->st);
        else
            peer_gather_put_kernel<float><<<vec_grid(h), kVecThreads, 0, h->stream>>>(h->pv, h->x, (long long)h->local_rows,
                                                                                     (long long)h->row_offset, buf, gseq, h->st);
        CK(cudaGetLastError());
        peer_gather_wait_kernel<<<1, 32, 0, h->stream>>>(h->pv, gseq, h->st);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(x, h->peer_base + h->pv.off_xg[buf], h->n * h->esz, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
        cudaError_t se = cudaStreamSynchronize(h->stream);
        if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "solution gather faulted on the device: %s", cudaGetErrorString(se));
        return check_device_error(h, h->h_st[2]); // 2: a peer rank never delivered its slice
    }
    if (h->comm_mode != kCommNccl) return h->fail(LAMCG_ERR_STATE, "get_solution over ranks needs an initialised communicator");
    if (h->local_rows)
        CK(cudaMemcpyAsync(h->x_full + h->row_offset * h->esz, h->x, h->local_rows * h->esz, cudaMemcpyDeviceToDevice, h->stream));
    int rc = allgather_vec(h, h->x_full);
    if (rc != LAMCG_OK) return rc;
    CK(cudaMemcpyAsync(x, h->x_full, h->n * h->esz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LAMCG_OK;
}

int lamcg_save_solution(lamcg_t *h, const char *path)
{
    if (!h || !path) return LAMCG_ERR_INVALID;
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no system");
    std::vector<char> x(h->n * h->esz);
    int rc = lamcg_get_solution(h, x.data());
    if (rc != LAMCG_OK) return rc;
    if (h->rank != 0) return LAMCG_OK; // only rank 0 saves (MPI_OMP.hpp:426)
    FILE *f = fopen(path, "wb");
    if (!f) return h->fail(LAMCG_ERR_IO, "Cannot open output file %s: %s", path, strerror(errno));
    const uint64_t hdr[2] = {(uint64_t)h->n, 1ull};
    bool ok = fwrite(hdr, sizeof hdr, 1, f) == 1 && fwrite(x.data(), h->esz, h->n, f) == h->n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return h->fail(LAMCG_ERR_IO, "short write to %s", path);
    return LAMCG_OK;
}

// ---- random SPD system generator (reference: challenge/main/random_spd_system.cpp) ----
namespace {

constexpr long long kGemmWsTile = 512 * 512; // largest M*N that takes the split-K path

int launch_gemm(lamcg *h, const double *A, const double *B, double *C, long long M, long long N, long long K, long long sai,
                long long sak, long long sbk, long long sbj, long long sci, long long scj, double alpha, double beta)
{
    if (M <= 0 || N <= 0) return LAMCG_OK;
    GemmArgs g{A, B, C, M, N, K, sai, sak, sbk, sbj, sci, scj, alpha, beta, 0, nullptr};
    // products with at least a 64 x 64 result go to the fp64 tensor cores (128 x 128 tiles); thinner ones (the bottom levels of
    // the Gram-Schmidt recursion: a handful of columns against n rows) stay on the 64 x 64 SIMT kernel
    const bool mma = M >= 64 && N >= 64 && !h->opt_spd_simt;
    const int tile = mma ? kMmaTM : 64;
    dim3 grid((unsigned)((N + tile - 1) / tile), (unsigned)((M + tile - 1) / tile));
    auto launch = [&](dim3 gr) {
        if (mma) gemm_f64_mma_kernel<<<gr, kMmaThreads, kMmaSmemBytes, h->stream>>>(g);
        else gemm_f64_kernel<<<gr, 256, 0, h->stream>>>(g);
    };
    // few output tiles but a long K (Q1^T Q2 of the lower and middle Gram-Schmidt levels): split along K over up to 2 CTAs per
    // SM worth of slices; slices are summed in slice order by a second kernel, so the result stays deterministic
    const long long tiles = (long long)grid.x * grid.y;
    if (h->gemm_ws && tiles < h->sm_count && K >= 2048 && M * N <= kGemmWsTile) {
        int slices = (int)std::min<long long>({128, K / 512, (long long)(2 * h->sm_count) / tiles});
        if (slices >= 2) {
            g.k_slice = ((K + slices - 1) / slices + 15) / 16 * 16;
            slices = (int)((K + g.k_slice - 1) / g.k_slice);
            g.ws = h->gemm_ws;
            grid.z = (unsigned)slices;
            launch(grid);
            CK(cudaGetLastError());
            gemm_splitk_reduce_kernel<<<(unsigned)std::min<long long>((M * N + 255) / 256, 1024), 256, 0, h->stream>>>(g, slices);
            CK(cudaGetLastError());
            return LAMCG_OK;
        }
    }
    launch(grid);
    CK(cudaGetLastError());
    return LAMCG_OK;
}

constexpr long long kSpdPanel = 32; // widest leaf handled by CholeskyQR2 instead of further recursion

// random_spd_system.cpp:41-62 — recursive block Gram-Schmidt on the columns [c0, c1) of the column-major Q (ld = n); leaves of
// up to kSpdPanel columns are orthonormalised by CholeskyQR2 (lamcg_spd.cuh: chol_inverse_kernel).  small: 2 * 32 * 32 doubles.
int gram_schmidt(lamcg *h, double *Q, long long n, long long c0, long long c1, double *buf, double *small)
{
    const long long cnt = c1 - c0;
    if (cnt == 1) {
        normalize_column_kernel<<<1, 256, 0, h->stream>>>(Q + c0 * n, n);
        CK(cudaGetLastError());
        return LAMCG_OK;
    }
    if (cnt <= kSpdPanel && !h->opt_spd_simt) {
        double *P = Q + c0 * n, *G = small, *Rinv = small + kSpdPanel * kSpdPanel;
        for (int pass = 0; pass < 2; ++pass) {
            int rc = launch_gemm(h, P, P, G, cnt, cnt, n, /*sai*/ n, /*sak*/ 1, /*sbk*/ 1, /*sbj*/ n, /*sci*/ 1, /*scj*/ cnt, 1.0, 0.0);
            if (rc != LAMCG_OK) return rc;
            chol_inverse_kernel<<<1, 32, 0, h->stream>>>(G, (int)cnt, Rinv);
            CK(cudaGetLastError());
            // in place: a 64-row tile of P is read completely (K = cnt <= 32, one column tile) before the CTA writes it back
            rc = launch_gemm(h, P, Rinv, P, n, cnt, cnt, /*sai*/ 1, /*sak*/ n, /*sbk*/ 1, /*sbj*/ cnt, /*sci*/ 1, /*scj*/ n, 1.0, 0.0);
            if (rc != LAMCG_OK) return rc;
        }
        return LAMCG_OK;
    }
    const long long mid = (c0 + c1) / 2, n1 = mid - c0, n2 = c1 - mid;
    int rc = gram_schmidt(h, Q, n, c0, mid, buf, small);
    if (rc != LAMCG_OK) return rc;
    const double *Q1 = Q + c0 * n;
    double *Q2 = Q + mid * n;
    // buf (n1 x n2, column-major, ld n1) = Q1^T Q2 ; Q2 -= Q1 buf
    rc = launch_gemm(h, Q1, Q2, buf, n1, n2, n, /*sai*/ n, /*sak*/ 1, /*sbk*/ 1, /*sbj*/ n, /*sci*/ 1, /*scj*/ n1, 1.0, 0.0);
    if (rc != LAMCG_OK) return rc;
    rc = launch_gemm(h, Q1, buf, Q2, n, n2, n1, /*sai*/ 1, /*sak*/ n, /*sbk*/ 1, /*sbj*/ n1, /*sci*/ 1, /*scj*/ n, -1.0, 1.0);
    if (rc != LAMCG_OK) return rc;
    return gram_schmidt(h, Q, n, mid, c1, buf, small);
}

// ---- glibc rand() restated (stdlib/random_r.c, TYPE_3): srand(seed) fills r[0..30] with the Park-Miller minimal standard
// generator, discards 310 outputs of v[k] = v[k-3] + v[k-31]; rand() returns v[k] >> 1.  State here: the 31 elements preceding
// the next output, oldest first.
struct GlibcState { uint32_t v[31]; };

GlibcState glibc_srand(unsigned seed)
{
    int32_t r[31];
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    // glibc keeps a ring with the front pointer at r[3] and the rear at r[0]: element k of the linear sequence is r[(k + 3) % 31]
    GlibcState s;
    for (int k = 0; k < 31; ++k) s.v[k] = (uint32_t)r[(k + 3) % 31];
    for (int i = 0; i < 310; ++i) { // one step: drop the oldest, append oldest + (element 28)
        const uint32_t nv = s.v[0] + s.v[28];
        for (int k = 0; k < 30; ++k) s.v[k] = s.v[k + 1];
        s.v[30] = nv;
    }
    return s;
}

// States at positions 0, chunk, 2 chunk, ... of the stream (nchunks x 31 words) by jump-ahead: M^chunk from repeated squaring of the
// companion matrix of the recurrence (arithmetic mod 2^32 is what uint32_t does).
std::vector<uint32_t> glibc_states(unsigned seed, long long chunk, long long nchunks)
{
    using Mat = std::vector<uint32_t>; // 31 x 31, row major
    auto mul = [](const Mat &A, const Mat &B) {
        Mat C(31 * 31, 0u);
        for (int i = 0; i < 31; ++i)
            for (int k = 0; k < 31; ++k) {
                const uint32_t a = A[i * 31 + k];
                if (a == 0u) continue;
                for (int j = 0; j < 31; ++j) C[i * 31 + j] += a * B[k * 31 + j];
            }
        return C;
    };
    Mat M(31 * 31, 0u), P(31 * 31, 0u);
    for (int i = 0; i < 30; ++i) M[i * 31 + i + 1] = 1u; // shift
    M[30 * 31 + 0] = 1u;                                  // new element = element 0 + element 28
    M[30 * 31 + 28] = 1u;
    for (int i = 0; i < 31; ++i) P[i * 31 + i] = 1u;
    for (long long e = chunk; e > 0; e >>= 1) {
        if (e & 1) P = mul(P, M);
        M = mul(M, M);
    }
    std::vector<uint32_t> out((size_t)nchunks * 31);
    GlibcState s = glibc_srand(seed);
    for (long long c = 0; c < nchunks; ++c) {
        memcpy(&out[(size_t)c * 31], s.v, sizeof s.v);
        GlibcState nx;
        for (int i = 0; i < 31; ++i) {
            uint32_t acc = 0u;
            for (int k = 0; k < 31; ++k) acc += P[i * 31 + k] * s.v[k];
            nx.v[i] = acc;
        }
        s = nx;
    }
    return out;
}

// out[0 .. count) <- the first `count` values 2*rand()/RAND_MAX - 1 after srand(seed), produced on the device
int device_random_fill(lamcg *h, double *out, long long count, int seed)
{
    const long long chunk = 31 * 128; // a multiple of the ring length
    const long long nchunks = (count + chunk - 1) / chunk;
    const std::vector<uint32_t> states = glibc_states((unsigned)seed, chunk, nchunks);
    unsigned int *d_states = nullptr;
    CK(cudaMalloc(&d_states, states.size() * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d_states, states.data(), states.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        glibc_rand_fill_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, h->stream>>>(out, count, d_states, chunk);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_states);
    if (e != cudaSuccess) return h->fail(LAMCG_ERR_CUDA, "device random fill failed: %s", cudaGetErrorString(e));
    return LAMCG_OK;
}

// random_spd_system.cpp:27-38 — glibc stream, column-major fill
void host_random_fill(double *out, size_t count, int seed)
{
    srand((unsigned)seed);
    for (size_t i = 0; i < count; ++i) out[i] = ((2.0 * rand()) / RAND_MAX) - 1.0;
}

} // namespace

int lamcg_random_spd_system(lamcg_t *h, size_t n, int seed)
{
    if (!h || n == 0) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (h->nranks != 1) return h->fail(LAMCG_ERR_INVALID, "the SPD generator runs on one rank (generate, save, then load row blocks)");
    if (h->dtype != 0) return h->fail(LAMCG_ERR_INVALID, "the SPD generator is fp64 only");
    if (mixed_storage(h))
        return h->fail(LAMCG_ERR_INVALID, "the SPD generator writes an fp64 matrix block (option matrix_f32 is on: generate with it off, save, then load)");
    int rc = alloc_system(h, n);
    if (rc != LAMCG_OK) return rc;
    CK(cudaFuncSetAttribute(gemm_f64_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmemBytes));
    double *Q = nullptr, *buf = nullptr, *d_dev = nullptr, *small = nullptr;
    auto cleanup = [&]() { cudaFree(Q); cudaFree(buf); cudaFree(d_dev); cudaFree(small); cudaFree(h->gemm_ws); h->gemm_ws = nullptr; };
    if (cudaMalloc(&h->gemm_ws, (size_t)128 * kGemmWsTile * sizeof(double)) != cudaSuccess) { cudaGetLastError(); h->gemm_ws = nullptr; }
    const size_t half = (n + 1) / 2;
    if (cudaMalloc(&Q, n * n * sizeof(double)) != cudaSuccess || cudaMalloc(&buf, (half * half + 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_dev, n * sizeof(double)) != cudaSuccess || cudaMalloc(&small, 2 * kSpdPanel * kSpdPanel * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        return h->fail(LAMCG_ERR_NOMEM, "device allocation for the %zu x %zu generator workspace failed", n, n);
    }
    const auto t_start = std::chrono::steady_clock::now();
    // Q <- U(-1,1): the glibc stream of srand(seed), generated on the device (column-major fill == stream order)
    rc = device_random_fill(h, Q, (long long)(n * n), seed);
    if (rc != LAMCG_OK) { cleanup(); return rc; }
    const auto t_fill = std::chrono::steady_clock::now();
    rc = gram_schmidt(h, Q, (long long)n, 0, (long long)n, buf, small);
    if (rc != LAMCG_OK) { cleanup(); return rc; }
    if (env_ll("spd_verbose", 0)) cudaStreamSynchronize(h->stream);
    const auto t_gs = std::chrono::steady_clock::now();
    {   // eigenvalues exp(3.5 U), seed - 10 ; scale column c by sqrt(D[c])
        std::vector<double> d(n);
        host_random_fill(d.data(), n, seed - 10);
        for (size_t i = 0; i < n; ++i) d[i] = std::exp(3.5 * d[i]);
        cudaError_t e = cudaMemcpyAsync(d_dev, d.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { cleanup(); return h->fail(LAMCG_ERR_CUDA, "upload of the eigenvalues failed: %s", cudaGetErrorString(e)); }
        scale_columns_kernel<<<std::min(h->sm_count * 8, kMaxGrid), 256, 0, h->stream>>>(Q, d_dev, (long long)n);
    }
    // A = Y Y^T into the padded row-major block (symmetric, so row-major == the reference's column-major file)
    if (h->lda != n) cudaMemsetAsync(h->A, 0, h->local_rows * h->lda * sizeof(double), h->stream);
    rc = launch_gemm(h, Q, Q, reinterpret_cast<double *>(h->A), (long long)n, (long long)n, (long long)n, /*sai*/ 1, /*sak*/ (long long)n, /*sbk*/ (long long)n,
                     /*sbj*/ 1, /*sci*/ (long long)h->lda, /*scj*/ 1, 1.0, 0.0);
    cudaError_t se = cudaStreamSynchronize(h->stream);
    const auto t_end = std::chrono::steady_clock::now();
    cleanup();
    if (rc != LAMCG_OK) return rc;
    if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "the SPD generator faulted on the device: %s", cudaGetErrorString(se));
    if (env_ll("spd_verbose", 0))
        fprintf(stderr, "lamcg_random_spd_system n=%zu: random fill %.3f s, Gram-Schmidt %.3f s, scale + Y Y^T %.3f s\n", n,
                std::chrono::duration<double>(t_fill - t_start).count(), std::chrono::duration<double>(t_gs - t_fill).count(),
                std::chrono::duration<double>(t_end - t_gs).count());
    h->has_matrix = true;
    std::vector<double> b(n);
    host_random_fill(b.data(), n, seed + 10); // random_spd_system.cpp:166
    return lamcg_set_rhs(h, b.data(), n);
}

int lamcg_save_system(lamcg_t *h, const char *matrix_path, const char *rhs_path)
{
    if (!h || !matrix_path || !rhs_path) return LAMCG_ERR_INVALID;
    if (h->nranks != 1) return h->fail(LAMCG_ERR_INVALID, "save_system runs on one rank");
    if (h->dtype != 0) return h->fail(LAMCG_ERR_INVALID, "save_system is fp64 only");
    if (mixed_storage(h)) return h->fail(LAMCG_ERR_INVALID, "save_system writes the fp64 matrix block (option matrix_f32 is on)");
    if (!h->has_matrix || !h->has_rhs) return h->fail(LAMCG_ERR_STATE, "no system to save");
    CK(cudaSetDevice(h->device));
    const size_t n = h->n;
    const uint64_t hdrA[2] = {(uint64_t)n, (uint64_t)n}, hdrb[2] = {(uint64_t)n, 1ull};
    FILE *f = fopen(matrix_path, "wb");
    if (!f) return h->fail(LAMCG_ERR_IO, "Cannot open output file %s: %s", matrix_path, strerror(errno));
    bool ok = fwrite(hdrA, sizeof hdrA, 1, f) == 1;
    const size_t chunk_rows = std::max<size_t>(1, ((size_t)64 << 20) / (n * sizeof(double)));
    std::vector<double> stage(chunk_rows * n);
    for (size_t r = 0; ok && r < n; r += chunk_rows) {
        const size_t nr = std::min(chunk_rows, n - r);
        cudaError_t e = cudaMemcpy2D(stage.data(), n * sizeof(double), h->A + r * h->lda * sizeof(double), h->lda * sizeof(double), n * sizeof(double), nr,
                                     cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { fclose(f); return h->fail(LAMCG_ERR_CUDA, "download of the matrix failed: %s", cudaGetErrorString(e)); }
        ok = fwrite(stage.data(), sizeof(double), nr * n, f) == nr * n;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return h->fail(LAMCG_ERR_IO, "short write to %s", matrix_path);
    std::vector<double> b(n);
    CK(cudaMemcpy(b.data(), h->b_full, n * sizeof(double), cudaMemcpyDeviceToHost));
    f = fopen(rhs_path, "wb");
    if (!f) return h->fail(LAMCG_ERR_IO, "Cannot open output file %s: %s", rhs_path, strerror(errno));
    ok = fwrite(hdrb, sizeof hdrb, 1, f) == 1 && fwrite(b.data(), sizeof(double), n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return h->fail(LAMCG_ERR_IO, "short write to %s", rhs_path);
    return LAMCG_OK;
}

// ---- measurement / test hooks ----------------------------------------------------------------
int lamcg_gemv(lamcg_t *h, const void *p, void *y_local, double *p_dot_y)
{
    if (!h || !p || !y_local) return LAMCG_ERR_INVALID;
    h->resumable = false; // the device state of the last solve no longer matches the system
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no matrix");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(p_ptr(h, 0), 0, h->lda * h->esz, h->stream));
    CK(cudaMemcpyAsync(p_ptr(h, 0), p, h->n * h->esz, cudaMemcpyDefault, h->stream));
    int rc = launch_gemv(h, 0);
    if (rc != LAMCG_OK) return rc;
    if (h->local_rows) CK(cudaMemcpyAsync(y_local, h->Ap, h->local_rows * h->esz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "GEMV faulted on the device: %s", cudaGetErrorString(se));
    if (p_dot_y) *p_dot_y = h->h_st[2].pAp_local;
    return check_device_error(h, h->h_st[2]);
}

int lamcg_vector_update_step(lamcg_t *h, size_t n, void *x, void *r, void *p, const void *Ap, double rr, double pAp, int fused,
                             double *alpha, double *rr_new, double *beta)
{
    if (!h || !x || !r || !p || !Ap || n == 0) return LAMCG_ERR_INVALID;
    if (h->nranks != 1) return h->fail(LAMCG_ERR_INVALID, "lamcg_vector_update_step is a single-rank test hook");
    h->resumable = false;
    CK(cudaSetDevice(h->device));
    const size_t bytes = n * h->esz;
    char *dv = nullptr; // x | r | Ap | p, private to this call: the hook works without a system
    CK(cudaMalloc(&dv, 4 * bytes));
    auto done = [&](int rc) { cudaFree(dv); return rc; };
    cudaError_t e = cudaMemcpyAsync(dv, x, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv + bytes, r, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv + 2 * bytes, Ap, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv + 3 * bytes, p, bytes, cudaMemcpyHostToDevice, h->stream);
    DevState s{}; // "entering the vector kernels of the first iteration"
    s.bb = rr;
    s.rr[0] = s.rr[1] = s.rr_final = rr;
    s.pAp_local = s.pAp = pAp;
    s.eps = 0.0;
    s.max_iters = 1 << 30; // not the last iteration: K3 updates p
    h->h_st[2] = s;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->st, &h->h_st[2], sizeof(DevState), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) return done(h->fail(LAMCG_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)));
    VecArgs v;
    v.st = h->st;
    v.pAp_src = &h->st->pAp_local;
    v.rrn_src = &h->st->rrn_local;
    v.x = dv;
    v.r = dv + bytes;
    v.Ap = dv + 2 * bytes;
    v.p_in = dv + 3 * bytes;
    v.p_out = dv + 3 * bytes;
    v.pv = no_peer();
    v.partials = h->partials + kMaxGrid;
    v.hist = nullptr;
    v.rows = (long long)n;
    v.row_offset = 0;
    v.par = 0;
    v.fused = fused ? 1 : 0;
    v.prof = 0;
    const int vg = (int)std::min<size_t>(std::max<size_t>((n + kVecThreads - 1) / kVecThreads, 1), (size_t)h->sm_count * 4); // vec_grid() of an n-row rank
    if (fused) {
        void *params[] = {&v};
        const void *fn = h->dtype == 0 ? (const void *)update_fused_kernel<double> : (const void *)update_fused_kernel<float>;
        e = cudaLaunchCooperativeKernel(fn, dim3(vg), dim3(kVecThreads), params, 0, h->stream);
    } else {
        if (h->dtype == 0) update_xr_kernel<double><<<vg, kVecThreads, 0, h->stream>>>(v);
        else update_xr_kernel<float><<<vg, kVecThreads, 0, h->stream>>>(v);
        if (h->dtype == 0) update_p_kernel<double><<<vg, kVecThreads, 0, h->stream>>>(v);
        else update_p_kernel<float><<<vg, kVecThreads, 0, h->stream>>>(v);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return done(h->fail(LAMCG_ERR_CUDA, "launch of the vector kernels failed: %s", cudaGetErrorString(e)));
    cudaMemcpyAsync(x, dv, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(r, dv + bytes, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(p, dv + 3 * bytes, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&h->h_st[2], h->st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) return done(h->fail(LAMCG_ERR_DEVICE, "the vector kernels faulted on the device: %s", cudaGetErrorString(se)));
    const DevState &o = h->h_st[2];
    if (alpha) *alpha = o.alpha_last;
    if (rr_new) *rr_new = o.rr[1];
    if (beta) *beta = o.beta_last;
    return done(check_device_error(h, o));
}

int lamcg_time_gemv(lamcg_t *h, int warmup, int reps, double *ms_per_launch)
{
    if (!h || reps <= 0 || !ms_per_launch) return LAMCG_ERR_INVALID;
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no matrix");
    CK(cudaSetDevice(h->device));
    for (int i = 0; i < warmup; ++i) {
        int rc = launch_gemv(h, 0);
        if (rc != LAMCG_OK) return rc;
    }
    CK(cudaEventRecord(h->ev_start, h->stream));
    for (int i = 0; i < reps; ++i) {
        int rc = launch_gemv(h, 0);
        if (rc != LAMCG_OK) return rc;
    }
    CK(cudaEventRecord(h->ev_stop, h->stream));
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess) return h->fail(LAMCG_ERR_DEVICE, "GEMV faulted on the device: %s", cudaGetErrorString(se));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
    *ms_per_launch = (double)ms / reps;
    return LAMCG_OK;
}

int lamcg_get_loop_profile(lamcg_t *h, long long *cycles_out, int capacity)
{
    if (!h || !cycles_out || capacity < 0) return LAMCG_ERR_INVALID;
    const int cnt = std::min(capacity, 8);
    for (int k = 0; k < cnt; ++k) cycles_out[k] = h->h_st[2].phase_cycles[k];
    return cnt;
}

int lamcg_time_stream_read(lamcg_t *h, int warmup, int reps, double *ms_per_pass, double *checksum)
{
    if (!h || reps <= 0 || !ms_per_pass) return LAMCG_ERR_INVALID;
    if (!h->has_matrix) return h->fail(LAMCG_ERR_STATE, "no matrix");
    CK(cudaSetDevice(h->device));
    const long long count2 = (long long)(h->local_rows * h->lda * h->asz / 16);
    const int grid = std::min(h->sm_count * 2, kMaxGrid);
    for (int i = 0; i < warmup; ++i) stream_read_kernel<<<grid, kStreamThreads, 0, h->stream>>>(reinterpret_cast<const double *>(h->A), count2, h->partials);
    CK(cudaEventRecord(h->ev_start, h->stream));
    for (int i = 0; i < reps; ++i) stream_read_kernel<<<grid, kStreamThreads, 0, h->stream>>>(reinterpret_cast<const double *>(h->A), count2, h->partials);
    CK(cudaEventRecord(h->ev_stop, h->stream));
    CK(cudaGetLastError());
    std::vector<double> part(grid);
    CK(cudaMemcpyAsync(part.data(), h->partials, grid * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
    *ms_per_pass = (double)ms / reps;
    if (checksum) {
        double s = 0.0;
        for (double v : part) s += v;
        *checksum = s;
    }
    return LAMCG_OK;
}

} // extern "C"
