// test_CG_single_GPU — positional-argument driver, drop-in for the reference executable of the same
// name (challenge/main/test/test_CG_single_GPU.cpp; also the calling convention of test_CG_CPU_OMP):
//   ./test_CG_single_GPU.out [matrix.bin [rhs.bin [sol.bin [max_iters [rel_error]]]]]
// defaults io/matrix.bin io/rhs.bin io/sol.bin 1000 1e-9; exit codes 1 (matrix), 2 (rhs), 6 (save).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "LAM.hpp"

int main(int argc, char **argv)
{
    const char *paths[3] = {"io/matrix.bin", "io/rhs.bin", "io/sol.bin"};
    int max_iters = 1000;
    double rel_error = 1e-9;
    for (int i = 0; i < 3; ++i)
        if (argc > i + 1) paths[i] = argv[i + 1];
    if (argc > 4) max_iters = std::atoi(argv[4]);
    if (argc > 5) rel_error = std::atof(argv[5]);

    std::printf("Usage: %s input_file_matrix.bin input_file_rhs.bin output_file_sol.bin max_iters rel_error\n", argv[0]);
    std::printf("All parameters are optional and have default values\n\n");
    std::printf("Command line arguments:\n");
    std::printf("  input_file_matrix: %s\n  input_file_rhs:    %s\n  output_file_sol:   %s\n", paths[0], paths[1], paths[2]);
    std::printf("  max_iters:         %d\n  rel_error:         %e\n\n", max_iters, rel_error);

    LAM::ConjugateGradient_B200<double> cg(0, 0, 1, LAM::Report::Text);
    if (!cg.ok()) return 1;

    std::printf("Reading matrix from file ...\n");
    if (!cg.load_matrix_from_file(paths[0])) {
        std::fprintf(stderr, "Failed to read matrix\n");
        return 1;
    }
    std::printf("Done\n\nReading right hand side from file ...\n");
    if (!cg.load_rhs_from_file(paths[1])) {
        std::fprintf(stderr, "Failed to read right hand side\n");
        return 2;
    }
    std::printf("Done\n\nSolving the system ...\n");

    const auto t0 = std::chrono::steady_clock::now();
    cg.solve(max_iters, rel_error);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Time elapsed using the B200 solver:%g s\n", secs);
    std::printf("Done\n\nWriting solution to file ...\n");
    if (!cg.save_result_to_file(paths[2])) {
        std::fprintf(stderr, "Failed to save solution\n");
        return 6;
    }
    std::printf("Done\n\nFinished successfully\n");
    return 0;
}
