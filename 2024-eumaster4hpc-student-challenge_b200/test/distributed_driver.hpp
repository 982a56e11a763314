// distributed_driver.hpp — the getopt driver shared by test_CG_MultiGPUS_CUDA_NCCL.cpp and test_CG_MultiGPUS_CUDA_MPI.cpp,
// drop-in for the reference's distributed executables (challenge/main/test/test_CG_CPU_MPI_OMP.cpp and
// test_CG_MultiGPUS_CUDA_{MPI,NCCL}.cpp, which differ from each other only in the solver class they instantiate):
//   -A <matrix> -b <rhs>   file mode          -s <n>   generate mode (exclusive with -A/-b)
//   -o <sol> -i <max_iters> -e <rel_error> -v -h      defaults io/*.bin, 10000, 1e-9
// Beyond the reference (its long generate-mode runs restart from x = 0):
//   -c <file>   write a checkpoint (x, r, p, scalars; one file per rank) if the solve stops on max_iters
//   -r <file>   restore that checkpoint after loading/generating the same system and run -i FURTHER iterations
// Ranks: one process per GPU, LAMCG_NGPUS=<P> in the environment plays the role of `srun -n P`.
// stdout (rank 0, not verbose) is the reference's one CSV line.  The including file sets LAMCG_DRIVER_PRINTS_COMM_INIT:
//   0 (…_MPI.out, like test_CPU_MPI_OMP.out):  n,ranks,threads,io_or_gen_s,avg_gemv_s,avg_iter_s,iters,rel_err,total_s          9 fields
//   1 (…_NCCL.out):                            n,ranks,threads,io_or_gen_s,comm_init_s,avg_gemv_s,avg_iter_s,iters,rel_err,total_s 10 fields
// The reference's NCCL class creates its communicator INSIDE solve(), prints the seconds that took right there
// (ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:306-334; e.g. TESTS/BEST_RESULTS:420 "10000,1,1,0.986,1.54201,...") and so counts them
// in total_s; this driver bootstraps its exchange (NVLink peer mapping, or NCCL with LAMCG_COMM=nccl) before the matrix exists, prints
// the same field in the same place and adds it to total_s.
// Quirks kept: iters = max_iters+1 when not converged, total_s whole seconds in generate mode (test_CG_CPU_MPI_OMP.cpp:176-178) and
// fractional in file mode (:87-92).  `threads` is the number of host threads that drive a rank's GPU: always 1 here (the reference
// prints omp_get_max_threads(), which its GPU job scripts pin to 1: every GPU row of TESTS/BEST_RESULTS has 1 in that column).
#pragma once

#ifndef LAMCG_DRIVER_PRINTS_COMM_INIT
#error "define LAMCG_DRIVER_PRINTS_COMM_INIT to 0 or 1 before including distributed_driver.hpp"
#endif

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <unistd.h>

#include "LAM.hpp"

namespace {

struct Args {
    const char *matrix = "io/matrix.bin";
    const char *rhs = "io/rhs.bin";
    const char *sol = "io/sol.bin";
    int max_iters = 10000;
    double rel_error = 1e-9;
    size_t n = 0;
    bool generate = false, load = false, verbose = false;
    const char *ckpt_out = nullptr, *ckpt_in = nullptr;
};

void usage(const char *prog)
{
    std::printf("Usage: %s [ (-A -b | -s) -o -e -i -h -v]\n", prog);
    std::printf("Options:\n");
    std::printf("  -A <file>       Read matrix from file\n");
    std::printf("  -b <file>       Read right hand side from file\n");
    std::printf("  -o <file>       Write solution to file\n");
    std::printf("  -i <int>        Maximum number of iterations\n");
    std::printf("  -e <float>      Relative error\n");
    std::printf("  -s <int>        Generate matrix of size n x n\n");
    std::printf("  -c <file>       Write a checkpoint if the solve stops on the iteration limit\n");
    std::printf("  -r <file>       Resume from a checkpoint (-i counts further iterations)\n");
    std::printf("  -v              Verbose mode\n");
    std::printf("  -h              Show this help message\n");
}

using Clock = std::chrono::steady_clock;
long long ms_since(Clock::time_point t0) { return std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - t0).count(); }

int run(LAM::RankWorld &world, const Args &a)
{
    const bool root = world.rank() == 0;
    const bool csv = root && !a.verbose;
    auto say = [&](const char *msg) { if (root && a.verbose) std::printf("%s", msg); };

    LAM::ConjugateGradient_B200<double> cg(world.rank(), world.rank(), world.size(), LAM::Report::Csv);
    if (!world.all_ok(cg.ok())) return 1; // all ranks leave together: nobody is left spinning in a bootstrap barrier
    if (a.ckpt_out) cg.set_option("loop_mode", 2); // the one-kernel loop for small n keeps p in shared memory: nothing to checkpoint
    size_t n_hint = a.n;
    if (!a.generate) { // file mode: the system size is the first header word (size_t rows)
        if (FILE *f = std::fopen(a.matrix, "rb")) {
            unsigned long long hdr[2] = {0, 0};
            if (std::fread(hdr, sizeof hdr, 1, f) == 1) n_hint = (size_t)hdr[0];
            std::fclose(f);
        }
    }
    const double comm_s = cg.init_comm(world, n_hint);
    if (!world.all_ok(cg.ok())) {
        if (root) std::fprintf(stderr, "Failed to initialise the multi-GPU exchange\n");
        return 1;
    }
    if (root && a.verbose) {
        std::printf("Command line arguments:\n");
        if (a.generate) std::printf("  rows:    %zu\n  cols:    %zu\n  size of the problem: %f GB\n", a.n, a.n, a.n * (double)a.n * 8 / 1024.0 / 1024.0 / 1024.0);
        else std::printf("  input_file_matrix: %s\n  input_file_rhs:    %s\n", a.matrix, a.rhs);
        std::printf("  output_file_sol:   %s\n  max_iters:         %d\n  rel_error:         %e\n", a.sol, a.max_iters, a.rel_error);
        std::printf("  Number of processes: %d\n  Number of threads: 1\n  communicator init: %f s\n\n", world.size(), comm_s);
    }

    auto t0 = Clock::now();
    const bool have_matrix = world.all_ok(a.generate ? cg.generate_matrix(a.n, a.n) : cg.load_matrix_from_file(a.matrix));
    const long long io_ms = ms_since(t0);
    if (!have_matrix) {
        if (root) std::fprintf(stderr, "Failed to read matrix\n");
        return 1;
    }
    if (csv) {
        std::cout << world.size() << "," << 1 << "," << io_ms / 1000.0 << ",";
        if (LAMCG_DRIVER_PRINTS_COMM_INIT) std::cout << comm_s << ",";
        std::cout << std::flush;
    }
    if (root && a.verbose) std::printf("Time elapsed for %s the matrix:%f s\n", a.generate ? "generating" : "reading", io_ms / 1000.0);

    const bool have_rhs = world.all_ok(a.generate ? cg.generate_rhs() : cg.load_rhs_from_file(a.rhs));
    if (!have_rhs) {
        if (root) std::fprintf(stderr, "Failed to read right hand side\n");
        return 2;
    }

    say("Solving the system ...\n");
    // ranks may finish a cold-cache ingest (or a checkpoint load) far apart: meet here, so that the first flag wait inside the
    // iteration measures the exchange and not the slowest rank's file system
    world.barrier();
    t0 = Clock::now();
    if (a.ckpt_in) {
        if (!world.all_ok(cg.load_checkpoint_from_file(a.ckpt_in))) {
            if (root) std::fprintf(stderr, "Failed to read checkpoint\n");
            return 3;
        }
        cg.resume(a.max_iters, a.rel_error);
    } else {
        cg.solve(a.max_iters, a.rel_error);
    }
    const long long cg_ms = ms_since(t0) + (LAMCG_DRIVER_PRINTS_COMM_INIT ? (long long)(comm_s * 1000.0) : 0);
    if (a.ckpt_out && !cg.last_result().converged && !cg.last_result().numerical_breakdown && !cg.save_checkpoint_to_file(a.ckpt_out)) {
        if (root) std::fprintf(stderr, "Failed to save checkpoint\n");
        return 7;
    }
    if (csv) {
        if (a.generate) std::cout << cg_ms / 1000; // whole seconds, as the reference prints in generate mode
        else std::cout << cg_ms / 1000.0;
        std::cout << std::flush;
    }
    if (root && a.verbose) {
        const lamcg_result &r = cg.last_result();
        const lamcg_info in = cg.info();
        const double it_s = r.iterations_run / r.solve_seconds;
        std::printf("\ncg_tot:%f s, %d iterations run, %.2f it/s, effective matrix stream %.1f GB/s per GPU\n", cg_ms / 1000.0,
                    r.iterations_run, it_s, 8.0 * in.local_rows * in.n * it_s / 1e9);
    }

    say("Writing solution to file ...\n");
    if (!cg.save_result_to_file(a.sol)) {
        if (root) std::fprintf(stderr, "Failed to save solution\n");
        return 6;
    }
    say("Done\n\nFinished successfully\n");
    return 0;
}

} // namespace

int main(int argc, char **argv)
{
    Args a;
    int opt;
    while ((opt = getopt(argc, argv, "hvA:b:o:i:e:s:c:r:")) != -1) {
        switch (opt) {
        case 'A':
        case 'b':
            if (a.generate) {
                std::fprintf(stderr, "Option -s cannot be used with -%c.\n", opt);
                return 1;
            }
            a.load = true;
            (opt == 'A' ? a.matrix : a.rhs) = optarg;
            break;
        case 's':
            if (a.load) {
                std::fprintf(stderr, "Option -A and -b cannot be used with -s.\n");
                return 1;
            }
            a.generate = true;
            a.n = (size_t)std::atoll(optarg);
            break;
        case 'o': a.sol = optarg; break;
        case 'i': a.max_iters = std::atoi(optarg); break;
        case 'e': a.rel_error = std::atof(optarg); break;
        case 'c': a.ckpt_out = optarg; break;
        case 'r': a.ckpt_in = optarg; break;
        case 'v': a.verbose = true; break;
        case 'h': usage(argv[0]); return 0;
        case '?':
            if (optopt == 'A' || optopt == 'b' || optopt == 'o' || optopt == 'i' || optopt == 'e' || optopt == 's' || optopt == 'c' || optopt == 'r')
                std::fprintf(stderr, "Option -%c requires an argument.\n", optopt);
            else if (std::isprint(optopt))
                std::fprintf(stderr, "Unknown option `-%c'.\n", optopt);
            else
                std::fprintf(stderr, "Unknown option character `\\x%x'.\n", optopt);
            return 1;
        default: std::abort();
        }
    }
    // Neither -s nor -A/-b: the reference falls through with an indeterminate exit code
    // (test_CG_CPU_MPI_OMP.cpp:281-291); the README documents file mode with the default paths.
    if (!a.generate) a.load = true;

    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0); // stdout carries exactly one CSV line; NCCL's banner goes to stderr
    LAM::RankWorld world = LAM::RankWorld::launch(0); // forks before any CUDA call
    int rc = run(world, a);
    if (world.rank() == 0) std::cout << std::endl;
    return world.finalize(rc);
}
