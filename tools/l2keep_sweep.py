#!/usr/bin/env python
"""Third-generation one-kernel loop (4096 <= n <= 16384): how much of A to load with the L2 evict-last policy.  usage: python tools/l2keep_sweep.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

for n in [int(x) for x in (sys.argv[1:] or ["10000", "8192", "12288", "16384", "5000", "4096"])]:
    s = lamcg_b200.Solver(0)
    s.generate_matrix(n, n)
    s.generate_rhs()
    s.set_option("loop_mode", 3)
    s.set_option("persist_variant", 3)
    iters = 1000
    line = []
    for mb in (0, 24, 40, 56, 64, 72, 80, 88, 96, 104, 112, 120):
        s.set_option("persist_l2_keep_mb", mb)
        s.solve(iters, 0.0)
        best = 0.0
        for _ in range(3):
            r = s.solve(iters, 0.0)
            best = max(best, r.iterations_run / r.solve_seconds)
        line.append(f"{mb} MB: {best:7.0f}")
    print(f"n={n:6d} ({8.0 * n * n / 1e6:6.0f} MB)  " + "  ".join(line) + " it/s", flush=True)
    s.close()
