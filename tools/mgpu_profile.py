#!/usr/bin/env python
"""Where a multi-GPU iteration spends its time outside the GEMV: per-rank SM cycles CTA 0 waits for the peers (p slices, p.Ap, r.r)
and works in the vector kernel (option loop_profile), next to the event-timed GEMV and the iteration time.
usage: torchrun --nproc-per-node P tools/mgpu_profile.py [n] [iters]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank, world, local = lamcg_b200.launch.world_from_env()
torch.cuda.set_device(local)
dist.init_process_group("gloo")
s = lamcg_b200.Solver(local, rank, world)
lamcg_b200.launch.bootstrap_comm(s, n=n, mode="peer", dist=dist)
s.generate_matrix(n, n)
s.generate_rhs()
s.set_option("loop_profile", 1)
out = {}
for name, opts in (("graph", {"loop_mode": 2, "time_gemv": 0}), ("stream+events", {"loop_mode": 0, "time_gemv": 1})):
    for k, v in opts.items():
        s.set_option(k, v)
    s.solve(iters, 1e-9)
    dist.barrier()
    r = s.solve(iters, 1e-9)
    prof = s.loop_profile()
    mhz = 1965.0
    out[name] = {"us_per_iteration": 1e6 * r.solve_seconds / r.iterations_run, "gemv_us": 1e6 * r.gemv_seconds / r.iterations_run,
                 "wait_pAp_us": prof[2] / iters / mhz, "xr_work_us": prof[3] / iters / mhz,
                 "wait_rr_us": prof[4] / iters / mhz, "p_phase_us": prof[5] / iters / mhz}
allout = [None] * world
dist.all_gather_object(allout, out)
if rank == 0:
    for name in out:
        print(name)
        for r_, o in enumerate(allout):
            print(f"  rank {r_}: " + "  ".join(f"{k}={v:8.2f}" for k, v in o[name].items()))
s.close()
dist.barrier()
dist.destroy_process_group()
