#!/usr/bin/env python
"""Fourth-generation one-kernel loop: how the gather is published and polled (replicas, owner vs staged stores, delay before
the first poll, back-off between poll rounds, polling load).  usage: python tools/gen4_poll_sweep.py [n]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = 2000
s = lamcg_b200.Solver(0)
s.random_spd_system(n, 42)
s.set_option("loop_mode", 3)
s.set_option("persist_variant", 4)
rows = []
for copies, publish, poll, delay, backoff in itertools.product((1, 2, 4), (0, 1), (0, 2), (0, 300, 600, 900, 1200), (0, 100, 400)):
    if backoff and delay not in (0, 600):
        continue
    for k, v in (("persist_ll_copies", copies), ("persist_publish", publish), ("persist_poll", poll), ("persist_poll_delay", delay),
                 ("persist_poll_backoff", backoff)):
        s.set_option(k, v)
    s.solve(iters, 0.0)
    rates = []
    for _ in range(3):
        r = s.solve(iters, 0.0)
        rates.append(r.iterations_run / r.solve_seconds)
    prof = [round(c / iters) for c in s.loop_profile()]
    rows.append((min(rates), copies, publish, poll, delay, backoff, rates, prof))
    print(f"copies={copies} publish={'staged' if publish else 'owner '} poll={'ld.cg' if poll else 'v4.u64'} delay={delay:4d} backoff={backoff:3d}: " +
          " ".join(f"{x / 1e3:6.1f}k" for x in rates) + f"  phases {prof}", flush=True)
rows.sort(reverse=True)
print("best by worst-of-3:")
for r in rows[:8]:
    print(f"  {r[0] / 1e3:6.1f}k  copies={r[1]} publish={r[2]} poll={r[3]} delay={r[4]} backoff={r[5]}  phases {r[7]}")
s.close()
