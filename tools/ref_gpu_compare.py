#!/usr/bin/env python
"""The reference's own GPU solver classes against this library ON THE SAME B200 (BASELINE configs[1]:
"n = 50000 on a single B200 vs test_CG_single_GPU"; configs[4]: file mode n = 2048).

The unmodified reference classes (oracle/_ref/ref_gpu_{single,multi}.out, built by `make -C oracle refgpu` from the
sources under /root/reference for sm_100) run in their own process.  Their solve() mallocs, uploads A from pageable
host memory, iterates and frees (ref: LAM/src/GPU/local/ConjugateGradient_GPU_CUDA.cu:225-316), so for each system the
harness solves twice on the same resident host matrix, with K0 and K1 iterations:
    seconds_per_iteration = (t(K1) - t(K0)) / (K1 - K0),    upload+setup = t(K0) - K0 * seconds_per_iteration.
This library then solves the same system with the same K1 and the two solutions are compared.
Measurement tool only — nothing here is on the product path.  Writes one JSON object."""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import lamcg_b200  # noqa: E402
import oracle  # noqa: E402
from oracle import fileformat, random_spd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gen", type=int, nargs="*", default=[10000, 20000, 50000], help="generate-mode sizes")
ap.add_argument("--file-n", type=int, nargs="*", default=[2048], help="file-mode (random SPD) sizes")
ap.add_argument("--multi-max-n", type=int, default=20000, help="largest generate-mode n for the multi variant (identical to single on 1 GPU)")
ap.add_argument("--k0", type=int, default=20)
ap.add_argument("--k1", type=int, default=520)
ap.add_argument("--variants", nargs="*", default=["single", "multi"])
ap.add_argument("--out", default=None)
a = ap.parse_args()


def host_mem_gb() -> float:
    with open("/proc/meminfo") as f:
        for line in f:
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    return 0.0


def slope(runs):
    """runs: harness dicts for K0, K1 (and optionally K0 again).  -> (s/iteration, setup seconds)."""
    t0s = [r["seconds"] for r in runs if r["max_iters"] == a.k0]
    t1s = [r["seconds"] for r in runs if r["max_iters"] == a.k1]
    per_it = (min(t1s) - min(t0s)) / (a.k1 - a.k0)
    return per_it, min(t0s) - a.k0 * per_it


rep = {"host_mem_available_gb": host_mem_gb(), "k0": a.k0, "k1": a.k1, "systems": []}
s = lamcg_b200.Solver(0)
tmp = tempfile.mkdtemp(prefix="lamcg_refgpu_")

for n in a.gen:
    need = 8.0 * n * n / 1e9
    if need * 1.15 > rep["host_mem_available_gb"]:
        rep["systems"].append({"mode": "generate", "n": n, "skipped": f"reference needs {need:.1f} GB of host memory"})
        continue
    row = {"mode": "generate", "n": n, "matrix_GB": need}
    s.generate_matrix(n, n)
    s.generate_rhs()
    s.solve(a.k0, 1e-9)
    r = s.solve(a.k1, 1e-9)
    x = s.solution()
    row["b200"] = {"iterations": r.iterations, "rel_residual": r.rel_residual, "iterations_per_s": r.iterations_run / r.solve_seconds,
                   "ms_per_iteration": 1e3 * r.solve_seconds / r.iterations_run, "gemv_ms": s.time_gemv(3, 20)}
    row["b200"]["gemv_GBps"] = 8.0 * n * n / row["b200"]["gemv_ms"] / 1e6
    o = oracle.cg_solve_generated(n, a.k1, 1e-9)
    row["b200"]["x_rel_l2_vs_cpu_oracle"] = float(np.linalg.norm(x - o.x) / np.linalg.norm(o.x))
    for v in a.variants:
        if not oracle.ref_gpu_available(v):
            row[v] = {"skipped": "oracle/_ref harness not built"}
            continue
        if v == "multi" and n > a.multi_max_n:
            continue
        xp = os.path.join(tmp, f"x_{v}_{n}.bin")
        try:
            runs = oracle.ref_gpu_solve(v, [a.k0, a.k1, a.k0, a.k1], 1e-9, n=n, x_path=xp, timeout=900)
        except Exception as e:  # the reference crashed or timed out: report, do not hide
            row[v] = {"failed": repr(e)[:300]}
            continue
        per_it, setup = slope(runs)
        xr = fileformat.read_vector(xp)
        row[v] = {"class": runs[0]["variant"], "devices": runs[0]["devices"], "iterations": runs[-1]["iters"], "rel_residual": runs[-1]["rel"],
                  "ms_per_iteration": 1e3 * per_it, "iterations_per_s": 1.0 / per_it, "setup_and_upload_s": setup,
                  "effective_GBps": 8.0 * n * n / per_it / 1e9, "solve_wall_s": [r_["seconds"] for r_ in runs],
                  "x_rel_l2_vs_b200": float(np.linalg.norm(xr - x) / np.linalg.norm(x)),
                  "x_rel_l2_vs_cpu_oracle": float(np.linalg.norm(xr - o.x) / np.linalg.norm(o.x)),
                  "cuda_error": runs[-1]["cuda_error"]}
        row[v]["b200_speedup_loop"] = row["b200"]["iterations_per_s"] * per_it
        os.unlink(xp)
    rep["systems"].append(row)
    print(json.dumps(row), flush=True)

for n in a.file_n:
    A, b = random_spd.random_spd_system(n, 42)
    Ap, bp = os.path.join(tmp, f"A_{n}.bin"), os.path.join(tmp, f"b_{n}.bin")
    fileformat.write_matrix(Ap, A)
    fileformat.write_matrix(bp, b)
    row = {"mode": "file", "n": n, "seed": 42}
    s.load_matrix(Ap)
    s.load_rhs(bp)
    s.solve(1000, 1e-9)
    r = s.solve(1000, 1e-9)
    x = s.solution()
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    row["cpu_oracle"] = {"iterations": o.iters, "rel_residual": o.rel}
    row["b200"] = {"iterations": r.iterations, "rel_residual": r.rel_residual, "iterations_per_s": r.iterations_run / r.solve_seconds,
                   "us_per_iteration": 1e6 * r.solve_seconds / r.iterations_run,
                   "x_rel_l2_vs_cpu_oracle": float(np.linalg.norm(x - o.x) / np.linalg.norm(o.x))}
    for v in a.variants:
        if not oracle.ref_gpu_available(v):
            row[v] = {"skipped": "oracle/_ref harness not built"}
            continue
        xp = os.path.join(tmp, f"x_{v}_{n}.bin")
        try:
            # converged run for parity (rel_err 1e-9), then fixed-length runs (rel_err 0) for the per-iteration time
            conv = oracle.ref_gpu_solve(v, [1000], 1e-9, A_path=Ap, b_path=bp, x_path=xp)[0]
            xr = fileformat.read_vector(xp)
            k0, k1 = a.k0, max(a.k1, 2020)
            runs = oracle.ref_gpu_solve(v, [k0, k1, k0, k1], 0.0, A_path=Ap, b_path=bp)
        except Exception as e:
            row[v] = {"failed": repr(e)[:300]}
            continue
        t0 = min(r_["seconds"] for r_ in runs if r_["max_iters"] == k0)
        t1 = min(r_["seconds"] for r_ in runs if r_["max_iters"] == k1)
        per_it = (t1 - t0) / (k1 - k0)
        row[v] = {"class": conv["variant"], "devices": conv["devices"], "iterations": conv["iters"], "rel_residual": conv["rel"],
                  "us_per_iteration": 1e6 * per_it, "iterations_per_s": 1.0 / per_it, "converged_solve_wall_s": conv["seconds"],
                  "x_rel_l2_vs_b200": float(np.linalg.norm(xr - x) / np.linalg.norm(x)),
                  "x_rel_l2_vs_cpu_oracle": float(np.linalg.norm(xr - o.x) / np.linalg.norm(o.x)),
                  "b200_speedup_loop": row["b200"]["iterations_per_s"] * per_it}
    rep["systems"].append(row)
    print(json.dumps(row), flush=True)

text = json.dumps(rep, indent=1)
if a.out:
    with open(a.out, "w") as f:
        f.write(text + "\n")
print(text)
