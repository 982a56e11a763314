#!/usr/bin/env python
"""Latency-regime benchmark (BASELINE config 5 class): it/s of the loop variants at small n."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

for n in [int(x) for x in (sys.argv[1:] or ["2048"])]:
    s = lamcg_b200.Solver(0)
    s.generate_matrix(n, n)
    s.generate_rhs()
    iters = 2000
    for name, opts in [("graph", {"loop_mode": 2}), ("stream", {"loop_mode": 1}),
                       ("persistent v3 (streaming sweep)", {"loop_mode": 3, "persist_variant": 3}),
                       ("persistent v4 (gathered Ap)", {"loop_mode": 3, "persist_variant": 4}),
                       ("persistent auto", {"loop_mode": 3, "persist_variant": 0})]:
        if n > 4096 and opts.get("persist_variant") == 4:
            continue
        for k, v in opts.items():
            s.set_option(k, v)
        s.solve(iters, 0.0)
        best = 0.0
        for _ in range(3):
            r = s.solve(iters, 0.0)
            best = max(best, r.iterations_run / r.solve_seconds)
        prof = s.loop_profile() if opts.get("loop_mode") == 3 else []
        print(f"n={n:6d} {name:46s} {best:10.0f} it/s  ({1e6 / best:6.2f} us/iteration)" + (f"  phase cycles/iteration {[round(c / iters) for c in prof]}" if prof else ""), flush=True)
    s.close()
