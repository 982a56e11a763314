#!/usr/bin/env python
"""Sensitivity of the fourth-generation one-kernel loop: replicas of the gathered-Ap array, polling load, matrix class
(generate-mode tridiagonal vs dense random SPD), repeated.  usage: python tools/gen4_probe.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

iters = 2000
for n in [int(x) for x in (sys.argv[1:] or ["2048"])]:
    for kind in ("generate", "spd"):
        s = lamcg_b200.Solver(0)
        if kind == "generate":
            s.generate_matrix(n, n)
            s.generate_rhs()
        else:
            s.random_spd_system(n, 42)
        s.set_option("loop_mode", 3)
        for variant, copies, poll in ((2, 0, 0), (4, 1, 0), (4, 2, 0), (4, 4, 0), (4, 1, 2), (4, 2, 2), (4, 1, 0), (2, 0, 0)):
            s.set_option("persist_variant", variant)
            s.set_option("persist_ll_copies", copies)
            s.set_option("persist_poll", poll)
            s.solve(iters, 0.0)
            rates = []
            for _ in range(4):
                r = s.solve(iters, 0.0)
                rates.append(r.iterations_run / r.solve_seconds)
            prof = s.loop_profile()
            print(f"n={n:5d} {kind:8s} gen{variant} copies={copies} poll={poll}: " + " ".join(f"{x / 1e3:6.1f}k" for x in rates) +
                  f"  phases {[round(c / iters) for c in prof]}", flush=True)
        s.close()
