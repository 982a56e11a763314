// RankWorld.hpp — "mpirun -n P" without MPI: one process per GPU on one box.
//
// The reference's distributed drivers get rank/size from MPI_Init / MPI_Comm_rank
// (challenge/main/test/test_CG_MultiGPUS_CUDA_NCCL.cpp:205-209) and exchange the NCCL id with
// MPI_Bcast (GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:320-327).  This image has
// no MPI, so the driver forks P-1 children BEFORE the first CUDA call; the few bytes of bootstrap
// data (NCCL unique id, peer-memory handles, exit codes) travel through an anonymous shared
// mapping guarded by a sense-reversing barrier.  Everything on the data path is on the GPUs.
#pragma once

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

namespace LAM {

class RankWorld {
public:
    static constexpr size_t kSlotBytes = 256;
    static constexpr int kMaxRanks = 16;

    // nranks <= 0: take LAMCG_NGPUS from the environment (default 1).
    static RankWorld launch(int nranks = 0)
    {
        if (nranks <= 0) {
            const char *e = std::getenv("LAMCG_NGPUS");
            nranks = e && *e ? std::atoi(e) : 1;
        }
        if (nranks < 1) nranks = 1;
        if (nranks > kMaxRanks) nranks = kMaxRanks;
        RankWorld w;
        w.size_ = nranks;
        w.rank_ = 0;
        if (nranks == 1) return w;
        void *mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (mem == MAP_FAILED) {
            std::perror("mmap");
            std::exit(111);
        }
        w.sh_ = new (mem) Shared();
        std::fflush(stdout);
        std::fflush(stderr);
        for (int r = 1; r < nranks; ++r) {
            pid_t pid = fork();
            if (pid < 0) {
                std::perror("fork");
                std::exit(111);
            }
            if (pid == 0) {
                w.rank_ = r;
                w.kids_.clear();
                return w;
            }
            w.kids_.push_back(pid);
        }
        return w;
    }

    int rank() const { return rank_; }
    int size() const { return size_; }

    void barrier()
    {
        if (size_ == 1) return;
        const int sense = !local_sense_;
        local_sense_ = sense;
        if (sh_->count.fetch_add(1) == size_ - 1) {
            sh_->count.store(0);
            sh_->sense.store(sense);
        } else {
            const auto t0 = std::chrono::steady_clock::now();
            while (sh_->sense.load() != sense) {
                std::this_thread::sleep_for(std::chrono::microseconds(50));
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(600)) {
                    std::fprintf(stderr, "[rank %d] barrier timeout: a peer rank died\n", rank_);
                    std::_Exit(112);
                }
            }
        }
    }

    // every rank contributes `bytes` (<= kSlotBytes); `all` receives size()*bytes in rank order
    void allgather(const void *mine, size_t bytes, void *all)
    {
        if (size_ == 1) {
            std::memcpy(all, mine, bytes);
            return;
        }
        std::memcpy(sh_->slots[rank_], mine, bytes);
        barrier();
        for (int r = 0; r < size_; ++r) std::memcpy(static_cast<char *>(all) + r * bytes, sh_->slots[r], bytes);
        barrier();
    }

    // true iff `mine` is true on every rank: every rank learns of a failure on any rank and they can leave together
    // instead of one rank returning early while the others wait out the barrier timeout
    bool all_ok(bool mine)
    {
        if (size_ == 1) return mine;
        const int flag = mine ? 1 : 0;
        int flags[kMaxRanks] = {0};
        allgather(&flag, sizeof flag, flags);
        for (int r = 0; r < size_; ++r)
            if (!flags[r]) return false;
        return true;
    }

    void bcast(void *buf, size_t bytes, int root)
    {
        if (size_ == 1) return;
        if (rank_ == root) std::memcpy(sh_->slots[root], buf, bytes);
        barrier();
        if (rank_ != root) std::memcpy(buf, sh_->slots[root], bytes);
        barrier();
    }

    // Children exit with their code; rank 0 reaps them and returns the worst code.
    int finalize(int code)
    {
        if (size_ == 1) return code;
        if (rank_ != 0) {
            std::fflush(stdout);
            std::fflush(stderr);
            std::_Exit(code);
        }
        int worst = code;
        for (pid_t k : kids_) {
            int st = 0;
            if (waitpid(k, &st, 0) > 0) {
                const int c = WIFEXITED(st) ? WEXITSTATUS(st) : 113;
                if (c != 0 && worst == 0) worst = c;
            }
        }
        return worst;
    }

private:
    struct Shared {
        std::atomic<int> count{0};
        std::atomic<int> sense{0};
        unsigned char slots[kMaxRanks][kSlotBytes];
    };
    Shared *sh_ = nullptr;
    int rank_ = 0, size_ = 1, local_sense_ = 0;
    std::vector<pid_t> kids_;
};

} // namespace LAM
