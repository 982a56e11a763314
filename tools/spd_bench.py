#!/usr/bin/env python
"""Time the GPU random SPD generator (lamcg_random_spd_system) with its products on the fp64 tensor cores (DMMA, default) and on
the SIMT kernel (option spd_simt), and compare the two matrices.  usage: python tools/spd_bench.py [n ...]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import lamcg_b200  # noqa: E402
from oracle import fileformat  # noqa: E402

for n in [int(x) for x in (sys.argv[1:] or ["2048", "8192", "16384"])]:
    mats = {}
    for name, simt in (("DMMA", 0), ("SIMT", 1)):
        s = lamcg_b200.Solver(0)
        s.set_option("spd_simt", simt)
        s.random_spd_system(min(n, 1024), 7)  # warm-up: module load, allocator
        t0 = time.perf_counter()
        s.random_spd_system(n, 42)
        dt = time.perf_counter() - t0
        print(f"n={n:6d} {name}: {dt:7.3f} s  ({4.0 * n ** 3 / dt / 1e12:6.2f} TFLOP/s counting 4 n^3)", flush=True)
        if n <= 8192:
            td = tempfile.mkdtemp()
            pa, pb = os.path.join(td, "A.bin"), os.path.join(td, "b.bin")
            s.save_system(pa, pb)
            mats[name] = fileformat.read_matrix(pa)
            os.remove(pa); os.remove(pb); os.rmdir(td)
        s.close()
    if len(mats) == 2:
        d = float(np.linalg.norm(mats["DMMA"] - mats["SIMT"]) / np.linalg.norm(mats["SIMT"]))
        print(f"n={n:6d} |A_dmma - A_simt| / |A_simt| = {d:.2e}", flush=True)
