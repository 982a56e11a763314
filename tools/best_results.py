#!/usr/bin/env python
"""Pick the best row per (section, n, ranks) out of sweep files — the job of the reference's
TESTS/results/clean.sh (sort the merged CSVs, keep the fastest run).  Rows are the reference CSV lines
n,ranks,threads,io_s,avg_gemv_s,avg_iter_s,iters,rel_err,total_s ; "best" = smallest avg_iter_s.
usage: python tools/best_results.py sweep1.txt [sweep2.txt ...]"""
import sys
from collections import OrderedDict

best = OrderedDict()
section = "?"
for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith("-"):
            name = line.strip("-")
            if name:
                section = name
            continue
        f = line.split(",")
        if len(f) not in (9, 10):
            continue
        try:
            n, ranks = int(f[0]), int(f[1])
            avg_iter = float(f[-4])
        except ValueError:
            continue
        key = (section, n, ranks)
        if key not in best or avg_iter < best[key][0]:
            best[key] = (avg_iter, line)
last = None
for (sec, n, ranks), (_, line) in sorted(best.items(), key=lambda kv: (kv[0][0], kv[0][1], kv[0][2])):
    if sec != last:
        print(f"-----------------{sec}-----------------")
        last = sec
    print(line)
