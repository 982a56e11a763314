#!/usr/bin/env python
"""Fourth-generation one-kernel loop: delay before the first poll x replicas, per system size.  usage: python tools/gen4_delay_sweep.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

iters = 2000
for n in [int(x) for x in (sys.argv[1:] or ["256", "512", "1024", "2048", "3000"])]:
    s = lamcg_b200.Solver(0)
    s.generate_matrix(n, n)
    s.generate_rhs()
    s.set_option("loop_mode", 3)
    s.set_option("persist_variant", 4)
    for copies in (1,):  # replicas of the gathered array were dropped after this sweep (profiles/r02_gen4_delay_sweep.log)
        line = []
        for delay in (0, 200, 400, 500, 600, 700, 800, 1000):
            s.set_option("persist_poll_delay", delay)
            s.solve(iters, 0.0)
            rates = []
            for _ in range(3):
                r = s.solve(iters, 0.0)
                rates.append(r.iterations_run / r.solve_seconds)
            line.append(f"{delay}:{min(rates) / 1e3:.0f}-{max(rates) / 1e3:.0f}k")
        print(f"n={n:5d} copies={copies}  " + "  ".join(line), flush=True)
    s.close()
