// Both instantiations the reference provides for its GPU classes (<double> and <float>) through the same
// header: generate mode n = 1000, 100 iterations; the residual after k iterations is ~ 1/(k sqrt(8n)).
#include <cmath>
#include <cstdio>
#include <vector>

#include "LAM.hpp"

template <typename T>
static bool run(const char *name)
{
    LAM::ConjugateGradient_B200<T> cg(0, 0, 1, LAM::Report::Quiet);
    if (!cg.ok() || !cg.generate_matrix(1000, 1000) || !cg.generate_rhs()) return false;
    cg.solve(100, (T)1e-9);
    const lamcg_result &r = cg.last_result();
    const double expect = 1.0 / (100.0 * std::sqrt(8000.0));
    const bool ok = r.iterations == 101 && std::fabs(r.rel_residual - expect) / expect < 2e-3;
    // caller-owned buffers of type T through the original challenge signature
    std::vector<T> A(16, T(0)), b(4, T(1)), x(4, T(0));
    for (int i = 0; i < 4; ++i) A[i * 4 + i] = T(2);
    const bool ok2 = cg.solve(A.data(), b.data(), x.data(), 4, 10, (T)1e-6) && std::fabs((double)x[0] - 0.5) < 1e-6;
    std::printf("%s %s (iterations %d, rel %.6e, x0 %.6f)\n", name, ok && ok2 ? "ok" : "FAILED", r.iterations, r.rel_residual, (double)x[0]);
    return ok && ok2;
}

int main()
{
    const bool a = run<double>("double");
    const bool b = run<float>("float");
    return a && b ? 0 : 1;
}
