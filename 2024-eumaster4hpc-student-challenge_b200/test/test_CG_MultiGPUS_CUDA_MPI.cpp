// test_CG_MultiGPUS_CUDA_MPI — stands in for the reference executable of that name (challenge/main/test/CMakeLists.txt:26-30,
// test/test_CG_MultiGPUS_CUDA_MPI.cpp): same options, same 9-field CSV line as test_CPU_MPI_OMP.out
// (test/test_CG_CPU_MPI_OMP.cpp:201-203).  Everything else: distributed_driver.hpp.
#define LAMCG_DRIVER_PRINTS_COMM_INIT 0
#include "distributed_driver.hpp"
