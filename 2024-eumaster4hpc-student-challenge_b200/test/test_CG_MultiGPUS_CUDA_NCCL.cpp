// test_CG_MultiGPUS_CUDA_NCCL — stands in for the reference executable of that name (challenge/main/test/CMakeLists.txt:20-24,
// test/test_CG_MultiGPUS_CUDA_NCCL.cpp): same options, same 10-field CSV line with the communicator-init seconds after io_s
// (GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:329-334).  Everything else: distributed_driver.hpp.
#define LAMCG_DRIVER_PRINTS_COMM_INIT 1
#include "distributed_driver.hpp"
