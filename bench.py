#!/usr/bin/env python
"""bench.py — CG iterations/s and GEMV HBM GB/s on the generate-mode SPD system (fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (B200, through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]       the reference's CPU solver on host cores
    torchrun ... bench.py --gpus N ...                              one rank per GPU (driver launches this)

A "step" is one solve() of ITERS CG iterations on the n x n generate-mode system (A = tridiag(1,2,1)
stored dense, b = 1, rel_error 1e-9 so no step stops early).  Workload = BASELINE.json configs[2]:
n = 100000 (80 GB of fp64), strong scaling: the same system row-partitioned over N GPUs.
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "cg_iterations_per_second_fp64_n100k"
UNIT = "iterations/s"


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_cpu_sample(n: int, iters: int, target_block_gb: float = 6.0):
    """Time the UNMODIFIED reference CPU solver (oracle/_ref, LAM::ConjugateGradient_CPU_MPI_OMP<double>,
    all host threads) on a bounded sample of the n x n generate-mode system: rank 0's 1/P row block
    (full-length rows) via the 1-rank MPI shim's LAMCG_SHIM_SIZE=P.  Returns (it/s extrapolated to the
    whole system = 1 / (P * t_iter_block), description dict)."""
    import oracle
    if not oracle.ref_available():
        # the oracle port of the same loop (kind "port"); dense generate + solve on the block is not
        # available there, so time the structured oracle and say so
        t0 = time.perf_counter()
        oracle.cg_solve_generated(n, iters, 1e-9)
        dt = time.perf_counter() - t0
        return iters / dt, {"kind": "port", "cores": oracle.num_threads(),
                            "sample": f"oracle port (structured O(n) matvec, NOT the dense stream), n={n}, {iters} iterations"}
    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    oracle.ref().ref_set_threads(ncpu)
    P = max(1, int(round(8.0 * n * n / (target_block_gb * 1e9))))
    os.environ["LAMCG_SHIM_SIZE"] = str(P)
    try:
        oracle.ref_gen_solve(n, 2, 1e-9)  # warm-up: page in, first touch
        r = oracle.ref_gen_solve(n, iters, 1e-9)
    finally:
        os.environ["LAMCG_SHIM_SIZE"] = "1"
    t_iter_block = r.seconds / iters
    rows = n // P
    value = 1.0 / (P * t_iter_block)
    cores = oracle.ref().ref_num_threads()
    desc = {"kind": "reference", "cores": cores, "cpu": host_cpu_model(),
            "sample": (f"unmodified reference test path (ConjugateGradient_CPU_MPI_OMP<double>::solve, -O3, OpenMP {cores} threads) "
                       f"on rows 0..{rows - 1} of the n={n} generate-mode system (1/{P} row block, full-length rows, "
                       f"{8.0 * rows * n / 1e9:.2f} GB), {iters} iterations, {t_iter_block * 1e3:.2f} ms per block iteration; "
                       f"value = 1/({P} x t_block_iter) = whole-system iterations/s on these cores"),
            "block_gemv_GBps": 8.0 * rows * n / t_iter_block / 1e9}
    return value, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n
    vals = []
    desc = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, desc = reference_cpu_sample(n, args.ref_iters)
        if i >= args.warmup:
            vals.append(v)
    total = time.perf_counter() - t_all
    value = sum(vals) / len(vals)
    desc = dict(desc, value=value, unit=UNIT)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / (args.warmup + args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"generate-mode SPD n={n} fp64 (BASELINE configs[2]), CPU reference on host cores", "n": n,
                       "iters_per_step": args.ref_iters},
            "cpu_baseline": desc,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_b200_arm(args):
    import numpy as np
    import torch
    import lamcg_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA library is the only implementation (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, iters = args.n, args.iters
    s = lamcg_b200.Solver(local_rank, rank, world)
    comm_note = None
    if world > 1:
        # default: fused NVLink peer-store exchange; if any rank cannot map its peers (no P2P/IPC on this
        # box) every rank falls back to the NCCL collectives together (both are GPU paths of this library)
        mode = args.comm
        if mode == "peer":
            ok = 1
            try:
                lamcg_b200.launch.bootstrap_comm(s, n=n, mode="peer", dist=dist)
            except lamcg_b200.LamcgError as e:
                ok, comm_note = 0, f"peer exchange unavailable ({e.message}); NCCL used"
            t = torch.tensor([ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                s.close()
                s = lamcg_b200.Solver(local_rank, rank, world)
                mode = "nccl"
                comm_note = comm_note or "peer exchange unavailable on another rank; NCCL used"
        if mode == "nccl":
            lamcg_b200.launch.bootstrap_comm(s, n=n, mode="nccl", dist=dist)
    if args.gemv_variant:
        s.set_option("gemv_variant", args.gemv_variant)
    t0 = time.perf_counter()
    s.generate_matrix(n, n)
    s.generate_rhs()
    gen_s = time.perf_counter() - t0
    info = s.info
    bytes_per_gemv = 8.0 * info.local_rows * n  # algorithmic bytes per GEMV launch on this rank (8 n^2 / P)

    # ---------------- device-resident timed region: value + roofline from the same K steps
    s.set_option("time_gemv", 1)  # CUDA events around every GEMV launch, on the solver's stream
    for _ in range(args.warmup):
        s.solve(iters, 1e-9)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    dev_s = gemv_s = 0.0
    launches = its = 0
    last = None
    for _ in range(args.steps):
        last = s.solve(iters, 1e-9)
        dev_s += last.solve_seconds
        gemv_s += last.gemv_seconds
        launches += last.kernel_launches
        its += last.iterations_run
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_s_max = max_over_ranks(dev_s)
    wall_max = max_over_ranks(wall)
    gemv_s_max = max_over_ranks(gemv_s)
    value = its / dev_s_max
    gemv_ms = 1e3 * gemv_s_max / its
    gemv_gbps = bytes_per_gemv / (gemv_ms * 1e-3) / 1e9

    # ---------------- same loop as a CUDA graph (no per-GEMV events), for the record
    s.set_option("time_gemv", 0)
    s.set_option("loop_mode", 2)
    s.solve(iters, 1e-9)
    barrier()
    g = s.solve(iters, 1e-9)
    graph_its = g.iterations_run / max_over_ranks(g.solve_seconds)

    # ---------------- end to end through the public API with HOST buffers
    b_host = torch.ones(n, dtype=torch.float64).pin_memory()
    x_host = torch.empty(n, dtype=torch.float64).pin_memory()
    s.set_rhs(b_host)
    s.solve(iters, 1e-9)
    barrier()
    e0 = time.perf_counter()
    e_its = 0
    for _ in range(args.steps):
        s.set_rhs(b_host)                 # H2D: this step's right-hand side from pinned host memory
        r = s.solve(iters, 1e-9)          # the reference-facing call
        s.solution(out=x_host)            # D2H: the step's result (x; all-gathered over ranks when N > 1)
        e_its += r.iterations_run
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - e0)
    e2e_value = e_its / e2e_wall

    stream_ms, _ = s.time_stream_read(1, 3)
    stream_gbps = 8.0 * info.local_rows * info.lda / (stream_ms * 1e-3) / 1e9

    if rank == 0:
        peak, peak_src = measured_peaks()
        traffic = None
        tpath = os.path.join(REPO, "profiles", "gemv_traffic.json")
        if os.path.exists(tpath):  # ncu capture of this kernel on this workload; only valid for the same n and rank count
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("n") == n and tj.get("ranks") == world:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"generate-mode SPD n={n} fp64 ({8.0 * n * n / 1e9:.0f} GB), {iters} CG iterations per step, "
                                   f"row-partitioned over {world} GPU(s) (BASELINE configs[2])",
                       "n": n, "iters_per_step": iters, "rel_error": 1e-9, "ranks": world,
                       "rows_per_gpu": int(info.local_rows), "gemv_variant": int(info.gemv_variant),
                       "gemv_grid": int(info.gemv_grid), "gemv_block": int(info.gemv_block), "gemv_smem": int(info.gemv_smem_bytes),
                       "comm": {0: "none", 1: "nccl", 2: "peer"}[int(info.comm_mode)], "comm_note": comm_note,
                       "loop": "stream launches with CUDA events around every GEMV (timed region); CUDA-graph loop reported in graph_iterations_per_s",
                       "l2": f"inputs larger than L2: {bytes_per_gemv / 1e9:.1f} GB streamed per GEMV per GPU vs 126 MB L2, no flush needed",
                       "generate_seconds": gen_s},
            "gemv_ms": gemv_ms, "gemv_GBps_per_gpu": gemv_gbps, "graph_iterations_per_s": graph_its,
            "wall_s_timed_region": wall_max, "rel_residual_after_step": last.rel_residual,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": gemv_gbps, "peak": peak, "unit": "GB/s", "frac": gemv_gbps / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "gemv (K1: Ap = A p + fused p.Ap)",
                         "algorithmic_bytes_per_launch": bytes_per_gemv, "avg_launch_ms": gemv_ms,
                         "read_only_stream_GBps": stream_gbps, "frac_of_read_only_stream": gemv_gbps / stream_gbps,
                         "frac_of_nominal_8000": gemv_gbps / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n + 64,
                    "note": "per step: b from pinned host memory (H2D), solve(), x back to pinned host memory (D2H); "
                            "A stays resident in HBM between steps as in the reference's load-once / generate-once flow"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, desc = reference_cpu_sample(n, args.ref_iters)
                line["cpu_baseline"] = dict(desc, value=v, unit=UNIT)
            except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=100, help="CG iterations per step (our arm)")
    ap.add_argument("--ref-iters", type=int, default=20, help="CG iterations per step of the CPU reference sample")
    ap.add_argument("--gemv-variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default=os.environ.get("LAMCG_COMM", "peer"), choices=["nccl", "peer"],
                    help="multi-GPU exchange: NCCL collectives or fused NVLink peer stores")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
