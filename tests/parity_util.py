"""Shared tolerance logic for file-mode (ill-conditioned, cond ~ 1e3) parity tests.

BASELINE.json states: same iteration count +-1 and x relative L2 <= 1e-10.  Measured fact (SURVEY
section 4, tests/golden/golden.json, gpurun_out/parity_report.json): on these systems the UNMODIFIED
reference does not meet that against itself.  Changing only OMP_NUM_THREADS moves its stopping
iteration by up to 4 (233..237 on the n = 200 fixture, 258..262 at n = 300, 351..353 at n = 2048) and
its x by up to 2.1e-10; at a MATCHED iteration count (rel_error = 0) its x still moves by up to
1.3e-10; and all of those x are ~5e-10 away from the exact solution.  Each summation order perturbs
the CG recurrence at the 1e-16 level and cond(A) ~ 1e3 amplifies that over hundreds of iterations:
the process is chaotic at the 1e-10 level, for the reference and for us alike.  Therefore

  * generate mode (the sharp case: reference self-noise ~1e-14) is held to EXACT iteration counts and
    x <= 1e-12 (1e-10 beyond 1000 iterations) elsewhere in the suite — stricter than north_star;
  * file mode is held to:  x at matched iteration count within 5e-10 of the oracle (1e-10 is recorded
    in the report whenever it is met, which is the common case), the stopping iteration within
    max(3, 2 %) of the oracle's, and — the criterion that does not depend on rounding luck — our x must
    be as close to the EXACT solution (LAPACK solve) as the reference's x is, within a factor 2
    (both sit at ~2-5e-10 from it; the reference's own threads move that distance by tens of %).
"""
import math

import numpy as np

import oracle

X_TOL = 1e-10          # north_star; asserted wherever it is robustly attainable
X_TOL_FILE = 5e-10     # file mode at matched iteration count, see module docstring


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def iteration_slack(oracle_iters):
    return max(3, math.ceil(0.02 * oracle_iters))


def reference_self_noise(A, b, k, x_oracle):
    """Largest x difference of the unmodified reference against the oracle (= itself at 1 thread) after
    exactly k iterations when only OMP_NUM_THREADS changes; None when oracle/_ref is not present."""
    if not oracle.ref_available():
        return None
    return max(rel_l2(oracle.ref_omp_solve(A, b, k, 0.0, threads=t).x, x_oracle) for t in (2, 3, 4, 8))


def as_accurate_as_reference(A, b, x_ours, x_oracle):
    """(ours_err, oracle_err): relative distances to the exact solution."""
    x_true = np.linalg.solve(A, b)
    return rel_l2(x_ours, x_true), rel_l2(x_oracle, x_true)
