#!/usr/bin/env python
"""One persistent-loop solve per generation with its phase timers: python tools/persist_profile.py [n] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 500
s = lamcg_b200.Solver(0)
s.generate_matrix(n, n)
s.generate_rhs()
s.set_option("loop_mode", 3)
names = ["p update", "GEMV", "row sums + p.Ap exchange", "alpha broadcast", "x/r update + r.r exchange", "beta broadcast"]
for gen in (1, 2, 3):
    if gen == 2 and n > 4096:
        continue
    s.set_option("persist_variant", gen)
    s.solve(iters, 0.0)
    r = s.solve(iters, 0.0)
    print(f"generation {gen}: n={n} iters={iters} {r.iterations_run / r.solve_seconds:.0f} it/s")
    prof = s.loop_profile()
    for nm, c in zip(names, prof):
        print(f"  {nm:28s} {c / r.iterations_run:9.0f} cycles/iteration")
    print(f"  total                        {sum(prof[:6]) / r.iterations_run:9.0f} cycles/iteration; wall {1e6 * r.solve_seconds / r.iterations_run:.2f} us/iteration")
s.close()

