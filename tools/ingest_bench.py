#!/usr/bin/env python
"""File-mode ingest: time lamcg_load_matrix (pread -> pinned double buffer -> async
2-D H2D copy, 64-bit sizes) against the reference's own loader (MPI-IO shim / fread, CSV column 4 of
oracle/_ref/test_CG_CPU_MPI_OMP.out) on the same file.  usage: python tools/ingest_bench.py [n]"""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import lamcg_b200  # noqa: E402
import oracle  # noqa: E402
from oracle import fileformat  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
td = tempfile.mkdtemp(prefix="lamcg_ingest_")
pa, pb, px = (os.path.join(td, f) for f in ("A.bin", "b.bin", "x.bin"))
A = oracle.generate_matrix(n)
fileformat.write_matrix(pa, A)
fileformat.write_matrix(pb, np.ones(n))
gb = 8.0 * n * n / 1e9
del A
s = lamcg_b200.Solver(0)
s.load_matrix(pa)  # warm: allocation + first touch of the device block
for chunk_mb in (8, 2, 32):
    for threads in (1, 2, 4, 8, 12, 16):
        s.set_option("ingest_threads", threads)
        s.set_option("ingest_chunk_bytes", chunk_mb << 20)
        s.load_matrix(pa)  # sizes the pinned pool for this shape (kept between loads)
        t0 = time.perf_counter()
        s.load_matrix(pa)
        dt = time.perf_counter() - t0
        print(f"lamcg_load_matrix n={n} ({gb:.2f} GB), {threads:2d} reader thread(s), {chunk_mb:2d} MB chunks ({s.info.ingest_chunks} chunks): "
              f"{dt:.3f} s = {gb / dt:.2f} GB/s into HBM (file in page cache)", flush=True)
s.set_option("ingest_threads", 8)
s.set_option("ingest_chunk_bytes", 8 << 20)
s.load_matrix(pa)
s.load_rhs(pb)
r = s.solve(15, 1e-9)
o = oracle.cg_solve_generated(n, 15, 1e-9)
x = s.solution()
print("solve after load: iterations", r.iterations, "oracle", o.iters, "x rel err", float(np.linalg.norm(x - o.x) / np.linalg.norm(o.x)))
s.close()
if os.path.exists(oracle.REF_TEST_MPI_OMP):
    out = subprocess.run([oracle.REF_TEST_MPI_OMP, "-A", pa, "-b", pb, "-o", px, "-i", "1"], capture_output=True, text=True).stdout
    f = out.strip().split(",")
    print(f"reference loader (test_CG_CPU_MPI_OMP.out, CSV col 4): {float(f[3]):.3f} s = {gb / float(f[3]):.2f} GB/s into host memory   [{out.strip()}]")
for f in (pa, pb, px):
    if os.path.exists(f):
        os.remove(f)
os.rmdir(td)
