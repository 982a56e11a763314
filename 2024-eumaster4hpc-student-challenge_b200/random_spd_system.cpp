// random_spd_system — GPU edition of the reference's input generator (challenge/main/random_spd_system.cpp,
// which needs Intel MKL and icpx: challenge/random_spd_system.sh).  Same command line, same file format,
// same random streams (glibc srand/rand: seed for Q, seed-10 for the eigenvalues, seed+10 for the rhs):
//   ./random_spd_system.out matrix_size output_file_matrix.bin output_file_rhs.bin random_seed
#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "lamcg.h"

int main(int argc, char **argv)
{
    std::printf("Usage: %s matrix_size output_file_matrix.bin output_file_rhs.bin random_seed\n", argv[0]);
    std::printf("All parameters are optional and have default values\n\n");
    size_t size = argc > 1 ? (size_t)std::atoll(argv[1]) : 10;
    const char *fm = argc > 2 ? argv[2] : "io/matrix.bin";
    const char *fr = argc > 3 ? argv[3] : "io/rhs.bin";
    const int seed = argc > 4 ? std::atoi(argv[4]) : (int)std::time(nullptr);
    std::printf("Command line arguments:\n  matrix_size:        %zu\n  output_file_matrix: %s\n  output_file_rhs:    %s\n  seed:               %d\n\n",
                size, fm, fr, seed);
    if ((long long)size <= 0) {
        std::fprintf(stderr, "Wrong argument value\n");
        return 1;
    }
    lamcg_t *h = nullptr;
    if (lamcg_create(&h, 0) != LAMCG_OK) {
        std::fprintf(stderr, "%s\n", lamcg_last_error(nullptr));
        return 1;
    }
    std::printf("Generating the matrix and the right hand side ...\n");
    if (lamcg_random_spd_system(h, size, seed) != LAMCG_OK) {
        std::fprintf(stderr, "%s\n", lamcg_last_error(h));
        return 1;
    }
    std::printf("Done\n\nWriting matrix and right hand side to file ...\n");
    if (lamcg_save_system(h, fm, fr) != LAMCG_OK) {
        std::fprintf(stderr, "%s\nFailed to save matrix\n", lamcg_last_error(h));
        return 2;
    }
    std::printf("Done\n\nFinished successfully\n");
    lamcg_destroy(h);
    return 0;
}
