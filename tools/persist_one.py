#!/usr/bin/env python
"""One solve through the persistent one-kernel loop (for ncu): python tools/persist_one.py <n> <iters> [persist_variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n, iters = int(sys.argv[1]), int(sys.argv[2])
s = lamcg_b200.Solver(0)
s.generate_matrix(n, n)
s.generate_rhs()
s.set_option("loop_mode", 3)
if len(sys.argv) > 3:
    s.set_option("persist_variant", int(sys.argv[3]))
r = s.solve(iters, 0.0)
print(n, iters, r.iterations_run, f"{r.iterations_run / r.solve_seconds:.0f} it/s")
s.close()
