// Abstract solver interface of the LAM library, B200 edition.
//
// Same contract as the reference's LAM::ConjugateGradient<FloatingType>
// (challenge/main/LAM/src/ConjugateGradient.hpp:9-28): a solver object owns its system, is filled
// through load_* (or generate_*, see the concrete class), solves with (max_iters, rel_error) and
// reports success as a bool.  A program written against the reference base class compiles
// unchanged against this one.
#pragma once

#include <type_traits>

namespace LAM {

template <typename FloatingType>
class ConjugateGradient {
    static_assert(std::is_floating_point<FloatingType>::value, "DataType must be floating point");

public:
    ConjugateGradient() = default;
    virtual ~ConjugateGradient() = default;

    // x0 = 0; iterate until sqrt(r.r / b.b) < rel_error or max_iters iterations; true iff converged.
    virtual bool solve(int max_iters, FloatingType rel_error) = 0;

    // Binary format: size_t rows, size_t cols, then rows*cols row-major values.
    virtual bool load_matrix_from_file(const char *filename) = 0;
    virtual bool load_rhs_from_file(const char *filename) = 0;
    virtual bool save_result_to_file(const char *filename) const = 0;
};

} // namespace LAM
