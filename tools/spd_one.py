#!/usr/bin/env python
"""One run of the GPU random SPD generator (for ncu: the DMMA GEMM kernel).  usage: python tools/spd_one.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = lamcg_b200.Solver(0)
s.random_spd_system(n, 42)
print("generated", n, flush=True)
s.close()
