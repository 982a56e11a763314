#!/usr/bin/env python
"""Launch the K1 kernel a few times on rank 0's row block of the n x n generate-mode system split over `ranks` ranks — the same
kernel, grid and bytes as inside the solve loop (minus the peer flag wait).  bench.py wraps this process in
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:lamcg_rowsweep_kernel` to measure roofline.traffic in the run.
usage: python tools/traffic_probe.py <n> <ranks>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamcg_b200  # noqa: E402

n, ranks = int(sys.argv[1]), int(sys.argv[2])
s = lamcg_b200.Solver(0, 0, ranks)
s.generate_matrix(n, n)
print("probe: rows", s.info.local_rows, "variant", s.info.gemv_variant, "ms per launch", s.time_gemv(1, 2), flush=True)
s.close()
