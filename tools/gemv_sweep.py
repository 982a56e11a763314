#!/usr/bin/env python
"""Time every GEMV kernel variant on this rank's block (CUDA events, lamcg_time_gemv) and the
read-only streaming ceiling.  Usage: python tools/gemv_sweep.py [n ...] [--variants 1,2,...] [--rows R]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="*", type=int, default=[100000])
ap.add_argument("--variants", default="36,46,32,42,11,2")
ap.add_argument("--ranks", type=int, default=1, help="emulate the row block of rank 0 of this many ranks")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rows_out = []
for n in a.n:
    s = lamcg_b200.Solver(0, 0, a.ranks)
    s.generate_matrix(n, n)
    info = s.info
    nbytes = 8.0 * info.local_rows * n
    ms, _ = s.time_stream_read(2, 5)
    print(f"n={n} rows={info.local_rows} lda={info.lda}  read-only stream: {ms:.3f} ms  {8.0 * info.local_rows * info.lda / ms / 1e6:.0f} GB/s", flush=True)
    rows_out.append({"n": n, "rows": info.local_rows, "variant": "stream_read", "ms": ms, "GBps": 8.0 * info.local_rows * info.lda / ms / 1e6})
    for v in [int(x) for x in a.variants.split(",")]:
        try:
            s.set_option("gemv_variant", v)
            i2 = s.info
            ms = s.time_gemv(3, a.reps)
            print(f"  variant {v:3d} grid={i2.gemv_grid:4d} block={i2.gemv_block:4d} smem={i2.gemv_smem_bytes:6d}: {ms:9.4f} ms  {nbytes / ms / 1e6:8.0f} GB/s", flush=True)
            rows_out.append({"n": n, "rows": info.local_rows, "variant": v, "grid": i2.gemv_grid, "ms": ms, "GBps": nbytes / ms / 1e6})
        except lamcg_b200.LamcgError as e:
            print(f"  variant {v}: {e}", flush=True)
    s.close()
if a.out:
    with open(a.out, "w") as f:
        json.dump(rows_out, f, indent=1)
