#!/usr/bin/env python
"""Regenerate tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Everything written here is small and committed; the GPU box never needs /root/reference.

What is pinned
  golden.json                generate mode via LAM::ConjugateGradient_CPU_MPI_OMP<double> (harness) and
                             via the reference CLI test_CG_CPU_MPI_OMP.out (CSV line), OMP_NUM_THREADS=1
  gen_x_n{8,1000,2048}.npy   the reference's private _x after solve()
  spd_n{64,200}_{A,b}.bin    random SPD systems in the reference file format (restated generator, seed 42)
  spd_n{64,200}_x.bin        solution files written by the reference CLI test_CG_CPU_OMP.out
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

import oracle  # noqa: E402
from oracle import fileformat, random_spd  # noqa: E402


def run_cli(args, threads=1):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    return subprocess.run(args, capture_output=True, text=True, env=env, check=True).stdout


def main():
    assert os.path.isdir("/root/reference/challenge/main"), "reference sources not present"
    oracle.build(ref=True)
    gold = {"generate_mode": [], "generate_mode_cli": [], "file_mode": [], "reference_result_dumps": []}

    # ---- generate mode through the class (harness), 1 thread => deterministic
    for n, max_iters in [(1, 100), (2, 100), (3, 100), (8, 100), (1000, 10000), (2048, 10000), (5001, 10000),
                         (2048, 15), (2048, 100), (2048, 1000), (4096, 10000)]:
        r = oracle.ref_gen_solve(n, max_iters, 1e-9, threads=1)
        o = oracle.cg_solve_generated(n, max_iters, 1e-9)
        entry = {"n": n, "max_iters": max_iters, "rel_error": 1e-9, "converged": r.converged, "iters": r.iters,
                 "rel_printed": r.rel, "x_norm2": float(np.linalg.norm(r.x)), "x_sum": float(r.x.sum()),
                 "oracle_rel": o.rel, "oracle_bit_identical_x": bool(np.array_equal(r.x, o.x))}
        gold["generate_mode"].append(entry)
        if (n, max_iters) in ((8, 100), (1000, 10000), (2048, 10000)):
            np.save(os.path.join(HERE, f"gen_x_n{n}.npy"), r.x)
        print(entry)

    # ---- generate mode through the reference CLI (the CSV contract)
    with tempfile.TemporaryDirectory() as td:
        for n, it in [(2048, 15), (10000, 15), (10000, 1000)]:
            out = run_cli([oracle.REF_TEST_MPI_OMP, "-s", str(n), "-i", str(it), "-e", "1e-9", "-o", os.path.join(td, "sol.bin")],
                          threads=os.cpu_count())
            f = out.strip().split(",")
            entry = {"n": n, "max_iters": it, "csv_fields": len(f), "n_field": int(f[0]), "ranks": int(f[1]),
                     "iters": int(f[6]), "rel_printed": float(f[7])}
            gold["generate_mode_cli"].append(entry)
            print(entry)

    # ---- file mode through the reference CLI test_CG_CPU_OMP.out
    for n in (64, 200):
        A, b = random_spd.random_spd_system(n, 42)
        pa, pb, px = (os.path.join(HERE, f"spd_n{n}_{k}.bin") for k in "Abx")
        fileformat.write_matrix(pa, A)
        fileformat.write_matrix(pb, b)
        out = run_cli([oracle.REF_TEST_OMP, pa, pb, px, "1000", "1e-9"], threads=1)
        m = re.search(r"Converged in (\d+) iterations, relative error is ([0-9.e+-]+)", out)
        rows, cols = fileformat.read_header(px)
        x = fileformat.read_vector(px)
        # normalise the reference's garbage upper header bits so the fixture is stable
        fileformat.write_matrix(px, x)
        o = oracle.cg_solve(A, b, 1000, 1e-9)
        # the reference's own reduction-order noise: the same unmodified solver, only OMP_NUM_THREADS varies
        spread = [oracle.ref_omp_solve(A, b, 1000, 1e-9, threads=t).iters for t in range(1, 9)]
        entry = {"n": n, "seed": 42, "iters": int(m.group(1)), "rel_printed": float(m.group(2)),
                 "iters_by_omp_threads_1_to_8": spread,
                 "cols_word_low32": cols & 0xFFFFFFFF, "oracle_iters": o.iters, "oracle_rel": o.rel,
                 "oracle_bit_identical_x": bool(np.array_equal(x, o.x)), "cond": float(np.linalg.cond(A))}
        gold["file_mode"].append(entry)
        print(entry)

    # ---- known answers printed in the reference's own result dumps (hardware independent)
    gold["reference_result_dumps"] = [
        {"source": "TESTS/BEST_RESULTS:173", "n": 80000, "max_iters": 15, "iters": 16, "rel_printed": 8.33333e-05},
        {"source": "TESTS/BEST_RESULTS:184", "n": 100000, "max_iters": 15, "iters": 16, "rel_printed": 7.45356e-05},
        {"source": "TESTS/BEST_RESULTS:214", "n": 200000, "max_iters": 15, "iters": 16, "rel_printed": 5.27046e-05},
        {"source": "TESTS/results/STRESS_TEST_GPU_MPI.txt:17", "n": 560000, "max_iters": 10, "iters": 11, "rel_printed": 4.72456e-05},
        {"source": "TESTS/results/STRESS_TEST_GPU_MPI.txt:18", "n": 560000, "max_iters": 3000, "iters": 3001, "rel_printed": 1.57485e-07},
        {"source": "TESTS/results/WEAK_SCALABILITY_GPU_MPI.txt:20", "n": 80000, "max_iters": 1000, "iters": 1001, "rel_printed": 1.25e-06},
    ]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
