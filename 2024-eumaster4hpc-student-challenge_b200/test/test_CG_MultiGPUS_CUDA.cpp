// test_CG_MultiGPUS_CUDA — positional-argument driver for several GPUs of one box, drop-in for the reference
// executable of the same name (challenge/main/test/test_CG_MultiGPUS_CUDA.cpp, single process driving all
// devices with cudaMemcpyPeerAsync).  Here: one forked rank per GPU (LAMCG_NGPUS=P, default 1), rows of A
// partitioned like the reference's distributed classes, exchange over NVLink peer stores (or NCCL).
//   ./test_CG_MultiGPUS_CUDA.out [matrix.bin [rhs.bin [sol.bin [max_iters [rel_error]]]]]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "LAM.hpp"

int main(int argc, char **argv)
{
    const char *matrix = argc > 1 ? argv[1] : "io/matrix.bin";
    const char *rhs = argc > 2 ? argv[2] : "io/rhs.bin";
    const char *sol = argc > 3 ? argv[3] : "io/sol.bin";
    const int max_iters = argc > 4 ? std::atoi(argv[4]) : 1000;
    const double rel_error = argc > 5 ? std::atof(argv[5]) : 1e-9;

    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    LAM::RankWorld world = LAM::RankWorld::launch(0); // forks before any CUDA call
    const bool root = world.rank() == 0;
    if (root) {
        std::printf("Usage: %s input_file_matrix.bin input_file_rhs.bin output_file_sol.bin max_iters rel_error\n", argv[0]);
        std::printf("All parameters are optional and have default values\n\nCommand line arguments:\n");
        std::printf("  input_file_matrix: %s\n  input_file_rhs:    %s\n  output_file_sol:   %s\n", matrix, rhs, sol);
        std::printf("  max_iters:         %d\n  rel_error:         %e\n  GPUs (ranks):      %d\n\n", max_iters, rel_error, world.size());
    }
    int rc = 0;
    {
        LAM::ConjugateGradient_B200<double> cg(world.rank(), world.rank(), world.size(), LAM::Report::Text);
        size_t n_hint = 0;
        if (FILE *f = std::fopen(matrix, "rb")) {
            unsigned long long hdr[2] = {0, 0};
            if (std::fread(hdr, sizeof hdr, 1, f) == 1) n_hint = (size_t)hdr[0];
            std::fclose(f);
        }
        // every outcome is agreed over the ranks (all_ok) so that they leave together
        if (!world.all_ok(cg.ok())) rc = 1;
        if (rc == 0) {
            cg.init_comm(world, n_hint);
            if (!world.all_ok(cg.ok())) {
                if (root) std::fprintf(stderr, "Failed to initialise the multi-GPU exchange\n");
                rc = 1;
            }
        }
        if (rc == 0 && !world.all_ok(cg.load_matrix_from_file(matrix))) {
            if (root) std::fprintf(stderr, "Failed to read matrix\n");
            rc = 1;
        }
        if (rc == 0 && !world.all_ok(cg.load_rhs_from_file(rhs))) {
            if (root) std::fprintf(stderr, "Failed to read right hand side\n");
            rc = 2;
        }
        if (rc == 0) {
            world.barrier(); // ranks may finish the file ingest far apart: meet before the first in-kernel flag wait
            const auto t0 = std::chrono::steady_clock::now();
            cg.solve(max_iters, rel_error);
            if (root) std::printf("Time elapsed using the B200 solver:%g s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            if (!cg.save_result_to_file(sol)) {
                if (root) std::fprintf(stderr, "Failed to save solution\n");
                rc = 6;
            } else if (root) {
                std::printf("Finished successfully\n");
            }
        }
    }
    return world.finalize(rc);
}
