#!/usr/bin/env python
"""Instruction mix of the K1 kernels in liblamcg.so, from `cuobjdump -sass` (works without a GPU).

    python tools/sass_summary.py [--out profiles/r02_sass_k1_summary.txt] [--listing profiles/r02_sass_rowsweep_default_mainloop.txt]

Per kernel: total instructions, global-load mnemonics with counts (the load WIDTH is the point: LDG.E.NA.128 for the
default row sweep, LDG.E.NA.ENL2.256 for the sm_100 256-bit shape), TMA (UBLKCP) and mbarrier (SYNCS) counts, FP64 ops
(DMUL/DADD unfused vs DFMA), spills (STL/LDL).  --listing also writes the main loop of the default kernel (the
longest backward-branch body that contains global loads).  tests/test_abi.py imports `kernels()` for its assertions.
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "liblamcg.so")
INSN = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)")


def kernels(lib: str = LIB) -> dict[str, list[tuple[int, str, str]]]:
    """{mangled name: [(address, opcode, full line), ...]}"""
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res: dict[str, list] = {}
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = res.setdefault(m.group(1), [])
            continue
        m = INSN.search(ln)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2), ln.rstrip()))
    return res


def mix(insns) -> collections.Counter:
    return collections.Counter(op for _, op, _ in insns)


def main_loop(insns):
    """Innermost streaming loop of the full-width pass: among the backward-branch bodies with at least 16 streaming (LDG...NA)
    loads, the one with the highest share of streaming loads among its global loads (8 rows per p vector in the main pass; the
    tail passes of 4 / 2 / 1 rows re-read p more often), shortest first; else the body with the most global loads."""
    spans = []
    for addr, op, ln in insns:
        if not op.startswith("BRA"):
            continue
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", ln)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr:
            continue
        body = [i for i in insns if tgt <= i[0] <= addr]
        spans.append((sum(1 for i in body if ".NA." in i[1]), sum(1 for i in body if i[1].startswith("LDG")), body))
    streaming = [sp for sp in spans if sp[0] >= 16]
    if streaming:
        return max(streaming, key=lambda sp: (round(sp[0] / sp[1], 3), -len(sp[2])))[2]
    return max(spans, key=lambda sp: sp[1])[2] if spans else []


def summarise(name, insns) -> str:
    c = mix(insns)
    pick = lambda pre: {k: v for k, v in sorted(c.items()) if k.startswith(pre)}
    loop = main_loop(insns)
    lc = mix(loop)
    lines = [f"{name}", f"  instructions: {len(insns)}   main loop: {len(loop)} instructions, "
             f"{sum(v for k, v in lc.items() if k.startswith('LDG'))} global loads, {lc.get('DMUL', 0)} DMUL, {lc.get('DADD', 0)} DADD, {lc.get('DFMA', 0)} DFMA"]
    for title, pre in (("global loads", "LDG"), ("global stores", "STG"), ("TMA bulk copies", "UBLKCP"), ("mbarrier", "SYNCS"),
                       ("shared loads", "LDS"), ("fp64", "D"), ("fp32", "F"), ("local (spill)", "STL"), ("local (spill)", "LDL"),
                       ("atomics / reductions", "ATOM"), ("atomics / reductions", "RED"), ("fences", "MEMBAR"), ("cluster", "UCGABAR")):
        d = pick(pre)
        if d:
            lines.append(f"  {title:22s} " + "  ".join(f"{k} x{v}" for k, v in d.items()))
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--listing")
    ap.add_argument("--match", default="rowsweep|warprows|tmaring|update_|cg_persistent|gemm_")
    a = ap.parse_args()
    ks = kernels()
    text = [f"# cuobjdump -sass {os.path.relpath(LIB, REPO)} — instruction mix per kernel (tools/sass_summary.py)", ""]
    for name in sorted(ks):
        if re.search(a.match, name):
            text.append(summarise(name, ks[name]))
            text.append("")
    blob = "\n".join(text)
    print(blob)
    if a.out:
        with open(os.path.join(REPO, a.out), "w") as f:
            f.write(blob)
    if a.listing:
        name = next(k for k in ks if "rowsweep_kernelIdLi8ELi4ELi512ELi1ELi16E" in k)
        with open(os.path.join(REPO, a.listing), "w") as f:
            f.write(f"# main loop (8 rows x 4 x 128-bit loads per thread, one 4096-column chunk) of {name}\n")
            f.write("\n".join(ln for _, _, ln in main_loop(ks[name])) + "\n")


if __name__ == "__main__":
    main()
