#!/usr/bin/env python
"""Single-GPU numbers for the BASELINE.json configs that are not the bench headline:
  config 1  generate n = 10 000, -i 1000 -e 1e-9      (reference prints 1001, 3.53553e-06)
  config 2  generate n = 50 000 (20 GB) on one B200
  config 5  file-mode n = 2048 (random_spd_system distribution), host buffers -> solve -> host x
Writes one JSON object (stdout and, with --out, a file)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lamcg_b200  # noqa: E402
import oracle  # noqa: E402
from oracle import random_spd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
a = ap.parse_args()
rep = {}
s = lamcg_b200.Solver(0)

for name, n, iters in (("config1_generate_n10000", 10000, 1000), ("config2_generate_n50000", 50000, 1000)):
    s.generate_matrix(n, n)
    s.generate_rhs()
    s.solve(20, 1e-9)
    r = s.solve(iters, 1e-9)
    gemv_ms = s.time_gemv(3, 20)
    info = s.info
    o = oracle.cg_solve_generated(n, iters, 1e-9)
    x = s.solution()
    rep[name] = {"n": n, "max_iters": iters, "iterations": r.iterations, "rel_residual": r.rel_residual,
                 "oracle_iterations": o.iters, "oracle_rel_residual": o.rel,
                 "x_rel_l2_vs_oracle": float(np.linalg.norm(x - o.x) / np.linalg.norm(o.x)),
                 "iterations_per_s": r.iterations_run / r.solve_seconds, "ms_per_iteration": 1e3 * r.solve_seconds / r.iterations_run,
                 "gemv_ms": gemv_ms, "gemv_GBps": 8.0 * n * n / gemv_ms / 1e6, "gemv_variant": int(info.gemv_variant)}
    print(name, rep[name], flush=True)

# config 5: file-mode class system, end to end from host buffers (A, b pinned) to host x
n = 2048
A, b = random_spd.random_spd_system(n, 42)
o = oracle.cg_solve(A, b, 1000, 1e-9)
At, bt = torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory()
xt = torch.empty(n, dtype=torch.float64).pin_memory()
cg = lamcg_b200.ConjugateGradient_B200(0, verbose=False)
cg.solve_system(At, bt, xt, 1000, 1e-9)
reps, t0 = 10, time.perf_counter()
for _ in range(reps):
    cg.solve_system(At, bt, xt, 1000, 1e-9)      # H2D of A (33.5 MB) + b, ~350 iterations, D2H of x
dt = (time.perf_counter() - t0) / reps
res = cg.last_result
x = xt.numpy()
rep["config5_file_n2048"] = {"n": n, "iterations": res.iterations, "oracle_iterations": o.iters,
                             "x_rel_l2_vs_oracle": float(np.linalg.norm(x - o.x) / np.linalg.norm(o.x)),
                             "loop_iterations_per_s": res.iterations_run / res.solve_seconds,
                             "e2e_seconds_per_solve_incl_matrix_upload": dt, "e2e_iterations_per_s": res.iterations_run / dt,
                             "h2d_bytes_per_solve": 8 * n * n + 8 * n, "d2h_bytes_per_solve": 8 * n}
print("config5_file_n2048", rep["config5_file_n2048"], flush=True)
if a.out:
    with open(a.out, "w") as f:
        json.dump(rep, f, indent=1)
