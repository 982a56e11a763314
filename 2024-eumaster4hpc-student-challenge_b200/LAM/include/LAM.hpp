// LAM.hpp — single include of the library, as in the reference (challenge/main/LAM/include/LAM.hpp).
// The reference pulls its CPU (OpenMP / MPI) and four CUDA variants here; this build has exactly one
// concrete solver, the B200-native one, and needs neither MPI nor OpenMP headers.
#pragma once

#include "../src/ConjugateGradient.hpp"
#include "../src/B200/RankWorld.hpp"
#include "../src/B200/ConjugateGradient_B200.hpp"

namespace LAM {
// Drop-in aliases: code written against the reference's GPU classes keeps compiling.
template <typename T> using ConjugateGradient_GPU_CUDA = ConjugateGradient_B200<T>;
template <typename T> using ConjugateGradient_MultiGPUS_CUDA = ConjugateGradient_B200<T>;
template <typename T> using ConjugateGradient_MultiGPUS_CUDA_MPI = ConjugateGradient_B200<T>;
template <typename T> using ConjugateGradient_MultiGPUS_CUDA_NCCL = ConjugateGradient_B200<T>;
} // namespace LAM
