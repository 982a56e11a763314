// ConjugateGradient_B200.hpp — the one concrete solver of this build.
//
// Public surface = the reference's distributed GPU classes
// (GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cuh:24-55: solve, load_matrix_from_file,
// load_rhs_from_file, save_result_to_file, generate_matrix, generate_rhs, get_num_rows,
// get_num_cols) plus the original challenge signature solve(A, b, x, size, max_iters, rel_error)
// that survives as a comment in test/test_CG_CPU_OMP.cpp:76-79.  Every method is a thin call into
// the C ABI of include/lamcg.h; all arithmetic happens in liblamcg.so on the GPU.  Behaviour kept
// from the reference: bool returns, messages on stderr, no exceptions, solve() == false when not
// converged, get_num_rows() == LOCAL rows, iteration count printed as max_iters+1 when not
// converged (MPI_OMP.hpp:125).
#pragma once

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <type_traits>

#include <unistd.h>

#include "lamcg.h"

#include "../ConjugateGradient.hpp"
#include "RankWorld.hpp"

namespace LAM {

enum class Report {
    Quiet, // print nothing
    Text,  // "Converged in %d iterations, ..." like ConjugateGradient_CPU_OMP::solve (OMP.hpp:80-90)
    Csv    // "n," from load/generate and "avg_gemv,avg_iter,iters,rel," from solve, like the
           // distributed classes (MPI_OMP.hpp:203-205,122-127); rank 0 only
};

template <typename FloatingType>
class ConjugateGradient_B200 : public ConjugateGradient<FloatingType> {
    static_assert(std::is_same<FloatingType, double>::value || std::is_same<FloatingType, float>::value,
                  "instantiated for double (the hot path) and float, like the reference's GPU classes");
    static constexpr int kDtype = std::is_same<FloatingType, double>::value ? 0 : 1;

public:
    explicit ConjugateGradient_B200(int device = 0, int rank = 0, int nranks = 1, Report report = Report::Text)
        : rank_(rank), nranks_(nranks), report_(report)
    {
        const int rc = lamcg_create_typed(&h_, device, rank, nranks, kDtype);
        if (rc != LAMCG_OK) {
            std::fprintf(stderr, "%s\n", lamcg_last_error(nullptr));
            h_ = nullptr;
        }
    }
    ~ConjugateGradient_B200() override { lamcg_destroy(h_); }
    ConjugateGradient_B200(const ConjugateGradient_B200 &) = delete;
    ConjugateGradient_B200 &operator=(const ConjugateGradient_B200 &) = delete;

    bool ok() const { return h_ != nullptr; }
    lamcg_t *handle() { return h_; }
    void set_report(Report r) { report_ = r; }
    bool set_option(const char *key, long long v) { return h_ && check(lamcg_set_option(h_, key, v)); }

    // Communicator bootstrap; returns seconds spent (the reference times and prints it, NCCL.cu:306-334).
    // Default: the fused NVLink peer-store exchange (needs the system size n up front because the
    // exchange buffers are exported before the matrix exists).  LAMCG_COMM=nccl, n == 0, or any rank
    // failing to map its peers makes ALL ranks use NCCL collectives instead.
    // Every step is agreed over all ranks (RankWorld::all_ok) before the next collective one, so a failure on one rank makes ALL
    // ranks give up together: afterwards ok() is false everywhere and nobody waits in a barrier for a rank that has left.
    double init_comm(RankWorld &world, size_t n = 0)
    {
        if (world.size() == 1) return 0.0;
        const auto t0 = std::chrono::steady_clock::now();
        auto give_up = [&]() {
            lamcg_destroy(h_);
            h_ = nullptr;
            return 0.0;
        };
        if (!world.all_ok(h_ != nullptr)) return give_up();
        const char *mode = std::getenv("LAMCG_COMM");
        bool peer = n > 0 && !(mode && std::string(mode) == "nccl");
        if (peer) {
            unsigned char mine[LAMCG_PEER_HANDLE_BYTES] = {0};
            std::vector<unsigned char> all((size_t)world.size() * LAMCG_PEER_HANDLE_BYTES);
            bool ok = lamcg_comm_peer_export(h_, n, mine) == LAMCG_OK;
            if (!ok) std::memset(mine, 0, sizeof mine);
            world.allgather(mine, sizeof mine, all.data());
            if (ok) ok = lamcg_comm_init_peer(h_, all.data()) == LAMCG_OK; // closes what it opened when it fails half way
            peer = world.all_ok(ok);
            if (!peer) { // start over with a clean handle so no half-mapped peer state survives
                const lamcg_info i = info();
                if (rank_ == 0) std::fprintf(stderr, "peer exchange unavailable (%s); using NCCL\n", lamcg_last_error(h_));
                lamcg_destroy(h_);
                h_ = nullptr;
                const bool again = lamcg_create_typed(&h_, i.device, rank_, nranks_, kDtype) == LAMCG_OK;
                if (!again) h_ = nullptr;
                if (!world.all_ok(again)) return give_up();
            }
        }
        if (!peer) {
            // NCCL prints its version banner on stdout when NCCL_DEBUG is set; stdout must carry exactly the
            // reference's CSV line, so fd 1 points at stderr while the communicator is created.
            std::fflush(stdout);
            std::cout.flush();
            const int saved_stdout = dup(1);
            dup2(2, 1);
            unsigned char id[LAMCG_NCCL_ID_BYTES] = {0};
            bool ok = true;
            if (world.rank() == 0 && lamcg_comm_nccl_unique_id(id) != LAMCG_OK) {
                std::fprintf(stderr, "%s\n", lamcg_last_error(nullptr));
                ok = false;
            }
            if (world.all_ok(ok)) {
                world.bcast(id, sizeof id, 0);
                ok = world.all_ok(check(lamcg_comm_init_nccl(h_, id)));
            } else {
                ok = false;
            }
            std::fflush(stdout);
            dup2(saved_stdout, 1);
            close(saved_stdout);
            if (!ok) return give_up();
        }
        world.barrier();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    // ---- reference interface -------------------------------------------------------------------
    bool solve(int max_iters, FloatingType rel_error) override
    {
        if (!h_ || !check(lamcg_solve(h_, max_iters, rel_error, &last_))) return false;
        report_solve(max_iters);
        return last_.converged != 0;
    }

    // ---- beyond the reference: continue a solve that ran out of iterations / checkpoint it -------
    // resume(m, eps) after solve(k, eps) that did not converge is bit-identical to solve(k + m, eps); reports totals.
    bool resume(int more_iters, FloatingType rel_error)
    {
        if (!h_ || !check(lamcg_solve_resume(h_, more_iters, rel_error, &last_))) return false;
        report_solve(last_.iterations - 1);
        return last_.converged != 0;
    }
    // One file per rank: <filename> on one rank, <filename>.rank<r>of<P> otherwise.
    bool save_checkpoint_to_file(const char *filename) const { return h_ && check(lamcg_checkpoint_save(h_, rank_path(filename).c_str())); }
    bool load_checkpoint_from_file(const char *filename) { return h_ && check(lamcg_checkpoint_load(h_, rank_path(filename).c_str())); }

    bool load_matrix_from_file(const char *filename) override
    {
        if (!h_ || !check(lamcg_load_matrix(h_, filename))) return false;
        print_n();
        return true;
    }
    bool load_rhs_from_file(const char *filename) override { return h_ && check(lamcg_load_rhs(h_, filename)); }
    bool save_result_to_file(const char *filename) const override { return h_ && check(lamcg_save_solution(h_, filename)); }

    virtual bool generate_matrix(size_t num_rows, size_t num_cols)
    {
        if (!h_ || !check(lamcg_generate_matrix(h_, num_rows, num_cols))) return false;
        print_n();
        return true;
    }
    virtual bool generate_rhs() { return h_ && check(lamcg_generate_rhs(h_)); }

    size_t get_num_rows() const { return info().local_rows; }
    size_t get_num_cols() const { return info().n; }

    // ---- original challenge signature: caller-owned buffers (host or device pointers) ------------
    bool solve(const FloatingType *A, const FloatingType *b, FloatingType *x, size_t size, int max_iters, FloatingType rel_error)
    {
        if (!h_ || !check(lamcg_set_matrix(h_, A, size, 0)) || !check(lamcg_set_rhs(h_, b, size))) return false;
        const bool converged = solve(max_iters, rel_error);
        if (!check(lamcg_get_solution(h_, x))) return false;
        return converged;
    }

    const lamcg_result &last_result() const { return last_; }
    lamcg_info info() const
    {
        lamcg_info i{};
        if (h_) lamcg_get_info(h_, &i);
        return i;
    }

private:
    // what the reference's solve() prints: CSV fields (MPI_OMP.hpp:122-127) or the human line (OMP.hpp:80-90)
    void report_solve(int max_iters)
    {
        double gemv_ms = 0.0;
        if (report_ == Report::Csv) lamcg_time_gemv(h_, 1, 3, &gemv_ms);
        if (rank_ != 0) return;
        if (report_ == Report::Csv) {
            const int its = last_.iterations_run > 0 ? last_.iterations_run : 1;
            std::cout << gemv_ms * 1e-3 << "," << last_.solve_seconds / its << "," << last_.iterations << "," << last_.rel_residual << ",";
        } else if (report_ == Report::Text) {
            if (last_.converged)
                std::printf("Converged in %d iterations, relative error is %e\n", last_.iterations, last_.rel_residual);
            else
                std::printf("Did not converge in %d iterations, relative error is %e\n", max_iters, last_.rel_residual);
        }
    }
    std::string rank_path(const char *filename) const
    {
        if (nranks_ == 1) return filename;
        return std::string(filename) + ".rank" + std::to_string(rank_) + "of" + std::to_string(nranks_);
    }
    bool check(int rc) const
    {
        if (rc == LAMCG_OK) return true;
        if (rank_ == 0) std::fprintf(stderr, "%s\n", lamcg_last_error(h_));
        return false;
    }
    void print_n() const
    {
        if (report_ == Report::Csv && rank_ == 0) std::cout << info().n << ",";
    }

    lamcg_t *h_ = nullptr;
    int rank_ = 0, nranks_ = 1;
    Report report_;
    lamcg_result last_{};
};

} // namespace LAM
