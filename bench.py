#!/usr/bin/env python
"""bench.py — CG iterations/s and GEMV HBM GB/s on the generate-mode SPD system (fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (B200, through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]       the reference's CPU solver on host cores
    torchrun ... bench.py --gpus N ...                              one rank per GPU (driver launches this)

A "step" is one solve() of ITERS CG iterations on the n x n generate-mode system (A = tridiag(1,2,1)
stored dense, b = 1, rel_error 1e-9 so no step stops early).  Workload = BASELINE.json configs[2]:
n = 100000 (80 GB of fp64), strong scaling: the same system row-partitioned over N GPUs.
One JSON line on stdout (rank 0).  Besides the contract keys (value, e2e, roofline, cpu_baseline, clocks, gpu_launches) the
line carries, all measured AFTER the timed region:
  parity         x and iteration count of the headline solve, of a remainder-row system (n = 100003: the last rank owns
                 n/P + n%P rows, MPI_OMP.hpp:175-184) and of a file-mode system ingested per rank through the chunked
                 multi-thread loader, each against the CPU oracle — at every N, so the scaling record carries multi-GPU parity;
  configs        BASELINE configs[0] (n = 10000), [1] (n = 50000), [4] (n = 2048 file mode) at N = 1 and [3] (n = 300000) at N = 8:
                 it/s, GEMV GB/s and parity;
  reference_gpu  ms/iteration of the UNMODIFIED reference GPU class (oracle/_ref/ref_gpu_single.out, GPU_CUDA.cu:225-316) on the
                 same B200 at n = 50000 (configs[1] names that comparison) and, host RAM permitting, at n = 100000 (N = 1 only);
  roofline.traffic  dram bytes of the K1 kernel measured NOW by ncu around a probe process that launches the same kernel on the
                 same row-block shape (falls back to the committed capture when ncu is unavailable).
See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "cg_iterations_per_second_fp64_n100k"
UNIT = "iterations/s"
K1_KERNEL_REGEX = "lamcg_rowsweep_kernel"


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_mem_gb() -> tuple[float, float]:
    """(MemTotal, MemAvailable) in GB."""
    tot = avail = 0.0
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemTotal"):
                    tot = int(line.split()[1]) / 1e6
                elif line.startswith("MemAvailable"):
                    avail = int(line.split()[1]) / 1e6
    except OSError:
        pass
    return tot, avail


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


def host_threads() -> int:
    """All the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------- CPU reference legs
def reference_cpu_block_sample(n: int, iters: int, target_block_gb: float = 6.0):
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
    oracle.ref().ref_set_threads(host_threads())
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
    tot, avail = host_mem_gb()
    desc = {"kind": "reference", "cores": cores, "cpu": host_cpu_model(), "host_mem_total_gb": tot, "host_mem_available_gb": avail,
            "same_system": False,
            "sample": (f"unmodified reference test path (ConjugateGradient_CPU_MPI_OMP<double>::solve, -O3, OpenMP {cores} threads) "
                       f"on rows 0..{rows - 1} of the n={n} generate-mode system (1/{P} row block, full-length rows, "
                       f"{8.0 * rows * n / 1e9:.2f} GB), {iters} iterations, {t_iter_block * 1e3:.2f} ms per block iteration; "
                       f"value = 1/({P} x t_block_iter) = whole-system iterations/s on these cores"),
            "block_gemv_GBps": 8.0 * rows * n / t_iter_block / 1e9}
    return value, desc


def run_reference_arm(args):
    """The reference's own CPU implementation on the box's host cores.  When the whole n x n system fits in host memory the
    UNMODIFIED reference generates it ONCE (80 GB at n = 100000) and every step is one solve() of --ref-iters iterations on it —
    the same system as the B200 arm, a bounded number of iterations per step; otherwise each step is the row-block sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    n = args.n
    tot, avail = host_mem_gb()
    need = 8.0 * n * n / 1e9
    full = oracle.ref_available() and not args.ref_block_sample and avail >= need * 1.12 + 4.0
    vals = []
    t_all = time.perf_counter()
    if full:
        cores = host_threads()
        sysm = None
        try:
            sysm = oracle.RefGenSystem(n, threads=cores)
        except MemoryError:
            full = False
    if full:
        its = max(1, args.ref_full_iters)
        secs = []
        last = None
        for i in range(args.warmup + args.steps):
            last = sysm.solve(its, 1e-9, want_x=(i == args.warmup + args.steps - 1))
            if i >= args.warmup:
                vals.append(its / last.seconds)
                secs.append(last.seconds)
        o = oracle.cg_solve_generated(n, its, 1e-9)
        import numpy as np
        x_err = float(np.linalg.norm(last.x - o.x) / np.linalg.norm(o.x))
        sysm.close()
        value = sum(vals) / len(vals)
        desc = {"kind": "reference", "cores": oracle.ref().ref_num_threads(), "cpu": host_cpu_model(), "host_mem_total_gb": tot,
                "host_mem_available_gb": avail, "same_system": True, "value": value, "unit": UNIT,
                "sample": (f"unmodified reference (ConjugateGradient_CPU_MPI_OMP<double>, -O3, OpenMP {oracle.ref().ref_num_threads()} threads) on the WHOLE "
                           f"n={n} generate-mode system ({need:.0f} GB in host memory, generated once in {sysm.gen_seconds:.1f} s), each step one solve() of "
                           f"{its} iterations ({1e3 * sum(secs) / len(secs) / its:.0f} ms per iteration); x vs oracle {x_err:.1e}"),
                "gemv_GBps": need * value, "iterations_per_step": its, "x_rel_l2_vs_oracle": x_err}
        iters_per_step = its
    else:
        desc = None
        for i in range(args.warmup + args.steps):
            v, desc = reference_cpu_block_sample(n, args.ref_iters)
            if i >= args.warmup:
                vals.append(v)
        value = sum(vals) / len(vals)
        desc = dict(desc, value=value, unit=UNIT)
        iters_per_step = args.ref_iters
    total = time.perf_counter() - t_all
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / (args.warmup + args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"generate-mode SPD n={n} fp64 ({need:.0f} GB), CPU reference on host cores (BASELINE configs[2])", "n": n,
                       "iters_per_step": iters_per_step, "rel_error": 1e-9, "whole_system": bool(full)},
            "cpu_baseline": desc,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- B200 arm
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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line; NCCL's version banner goes to stderr
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

    def bcast_obj(obj):
        if dist is None:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm_state = {"mode": args.comm, "note": None}

    def make_solver(n: int):
        """A ranked solver with its exchange bootstrapped for system size n.  Default: fused NVLink peer-store exchange; if any
        rank cannot map its peers (no P2P/IPC on this box) every rank falls back to the NCCL collectives together."""
        s = lamcg_b200.Solver(local_rank, rank, world)
        if world > 1:
            mode = comm_state["mode"]
            if mode == "peer":
                ok = 1
                try:
                    lamcg_b200.launch.bootstrap_comm(s, n=n, mode="peer", dist=dist)
                except lamcg_b200.LamcgError as e:
                    ok, comm_state["note"] = 0, f"peer exchange unavailable ({e.message}); NCCL used"
                t = torch.tensor([ok], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 0:
                    s.close()
                    s = lamcg_b200.Solver(local_rank, rank, world)
                    mode = comm_state["mode"] = "nccl"
                    comm_state["note"] = comm_state["note"] or "peer exchange unavailable on another rank; NCCL used"
            if mode == "nccl":
                lamcg_b200.launch.bootstrap_comm(s, n=n, mode="nccl", dist=dist)
        if args.gemv_variant:
            s.set_option("gemv_variant", args.gemv_variant)
        return s

    def rel_l2(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    n, iters = args.n, args.iters
    s = make_solver(n)
    t0 = time.perf_counter()
    s.generate_matrix(n, n)
    s.generate_rhs()
    gen_s = time.perf_counter() - t0
    info = s.info
    local_rows, lda = int(info.local_rows), int(info.lda)
    gemv_variant, gemv_grid, gemv_block, gemv_smem = int(info.gemv_variant), int(info.gemv_grid), int(info.gemv_block), int(info.gemv_smem_bytes)
    bytes_per_gemv = 8.0 * local_rows * n  # algorithmic bytes per GEMV launch on this rank (8 n^2 / P)

    # ---------------- device-resident timed region: value + roofline from the same K steps
    # the library's default engine at this size (CUDA-graph loop, 16 iterations per graph launch) with CUDA events recorded INSIDE the
    # captured chunks (external event-record nodes on the solver's stream) around ONE GEMV launch per chunk: an event node between
    # two kernels costs ~7 us on a busy 8-GPU box, i.e. 1 % of an iteration at N = 8 with events around every GEMV (measured:
    # 695 vs 702 it/s); every GEMV launch is the same work, so a sample of 1 in 16 inside the timed region times the kernel
    s.set_option("loop_mode", 2)
    s.set_option("time_gemv", 2)
    for _ in range(args.warmup):
        s.solve(iters, 1e-9)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    dev_s = gemv_s = 0.0
    launches = its = gemv_timed = 0
    last = None
    for _ in range(args.steps):
        last = s.solve(iters, 1e-9)
        dev_s += last.solve_seconds
        gemv_s += last.gemv_seconds
        gemv_timed += last.gemv_launches_timed
        launches += last.kernel_launches
        its += last.iterations_run
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_s_max = max_over_ranks(dev_s)
    wall_max = max_over_ranks(wall)
    gemv_s_max = max_over_ranks(gemv_s)
    value = its / dev_s_max
    gemv_ms = 1e3 * gemv_s_max / max(gemv_timed, 1)  # average duration of the timed GEMV launches (max over ranks of the sum)
    gemv_gbps = bytes_per_gemv / (gemv_ms * 1e-3) / 1e9

    # ---------------- parity of the headline solve (x of the last timed step vs the CPU oracle), every N
    x_head = s.solution()  # collective
    parity = {}
    if rank == 0:
        import oracle
        o = oracle.cg_solve_generated(n, iters, 1e-9)
        parity["headline"] = {"n": n, "ranks": world, "iterations": int(last.iterations), "oracle_iterations": int(o.iters),
                              "rel_residual": last.rel_residual, "oracle_rel_residual": o.rel, "x_rel_l2_vs_oracle": rel_l2(x_head, o.x),
                              "bound": 1e-12}
        parity["headline"]["ok"] = bool(last.iterations == o.iters and parity["headline"]["x_rel_l2_vs_oracle"] <= 1e-12)

    # ---------------- end to end through the public API with HOST buffers, same loop engine as the timed region
    b_host = torch.ones(n, dtype=torch.float64).pin_memory()
    x_host = torch.empty(n, dtype=torch.float64).pin_memory()
    s.set_rhs(b_host)
    s.solve(iters, 1e-9)

    def e2e_pass():
        barrier()
        e0 = time.perf_counter()
        e_its = 0
        for _ in range(args.steps):
            s.set_rhs(b_host)                 # H2D: this step's right-hand side from pinned host memory
            r = s.solve(iters, 1e-9)          # the reference-facing call
            s.solution(out=x_host)            # D2H: the step's result (x; all-gathered over ranks when N > 1)
            e_its += r.iterations_run
        barrier()
        return e_its / max_over_ranks(time.perf_counter() - e0)

    e2e_value = e2e_pass()

    # ---------------- for the record: the same graph loop without the per-GEMV events, and plain stream launches with events
    s.set_option("time_gemv", 0)
    s.solve(iters, 1e-9)
    barrier()
    g = s.solve(iters, 1e-9)
    graph_its = g.iterations_run / max_over_ranks(g.solve_seconds)
    s.set_option("loop_mode", 1)
    s.set_option("time_gemv", 1)
    s.solve(iters, 1e-9)
    barrier()
    g = s.solve(iters, 1e-9)
    stream_its = g.iterations_run / max_over_ranks(g.solve_seconds)
    stream_gemv_ms = 1e3 * max_over_ranks(g.gemv_seconds) / max(g.gemv_launches_timed, 1)

    stream_ms, _ = s.time_stream_read(1, 3)
    stream_gbps = 8.0 * local_rows * lda / (stream_ms * 1e-3) / 1e9
    comm_mode = {0: "none", 1: "nccl", 2: "peer"}[int(s.info.comm_mode)]
    s.close()
    del s
    barrier()

    # ---------------- more parity, after the timed region: remainder rows and per-rank file ingest, every N
    def generate_case(nn: int, k: int, bound: float, time_gemv_reps: int = 0):
        """Generate-mode solve of size nn, k iterations, through a fresh ranked solver; rank 0 returns the parity dict."""
        sv = make_solver(nn)
        sv.generate_matrix(nn, nn)
        sv.generate_rhs()
        sv.solve(min(k, 20), 1e-9)
        barrier()
        r = sv.solve(k, 1e-9)
        secs = max_over_ranks(r.solve_seconds)
        x = sv.solution()
        inf = sv.info
        out = {"n": nn, "ranks": world, "rows_on_this_rank": int(inf.local_rows), "rows_on_last_rank": nn // world + nn % world,
               "iterations": int(r.iterations), "rel_residual": r.rel_residual,
               "iterations_per_s": r.iterations_run / secs, "ms_per_iteration": 1e3 * secs / max(r.iterations_run, 1), "kernel_launches": int(r.kernel_launches)}
        if time_gemv_reps:
            ms = max_over_ranks(sv.time_gemv(3, time_gemv_reps))
            out.update(gemv_ms=ms, gemv_GBps_per_gpu=8.0 * int(inf.local_rows) * nn / ms / 1e6, gemv_variant=int(inf.gemv_variant))
        sv.close()
        if rank == 0:
            import oracle
            o = oracle.cg_solve_generated(nn, k, 1e-9)
            out.update(oracle_iterations=int(o.iters), oracle_rel_residual=o.rel, x_rel_l2_vs_oracle=rel_l2(x, o.x), bound=bound)
            out["ok"] = bool(out["iterations"] == o.iters and out["x_rel_l2_vs_oracle"] <= bound)
        return out

    def file_case(nn: int, k: int, chunk_bytes: int, threads: int):
        """File mode over ranks: rank 0 writes a dense SPD system in the reference format, every rank ingests ITS row block through
        the chunked multi-thread loader (small chunks: many chunks per rank, slot reuse), k iterations, x vs the dense oracle."""
        import oracle
        from oracle import fileformat
        tmp = bcast_obj(tempfile.mkdtemp(prefix="lamcg_bench_") if rank == 0 else None)
        pa, pb = os.path.join(tmp, "A.bin"), os.path.join(tmp, "b.bin")
        A = b = None
        if rank == 0:
            rng = np.random.default_rng(2024)
            B = rng.standard_normal((nn, nn))
            A = B @ B.T / nn + np.eye(nn)
            b = rng.standard_normal(nn)
            fileformat.write_matrix(pa, A)
            fileformat.write_matrix(pb, b)
        barrier()
        sv = make_solver(nn)
        sv.set_option("ingest_chunk_bytes", chunk_bytes)
        sv.set_option("ingest_threads", threads)
        t0 = time.perf_counter()
        sv.load_matrix(pa)
        load_s = max_over_ranks(time.perf_counter() - t0)
        sv.load_rhs(pb)
        inf = sv.info
        barrier()
        r = sv.solve(k, 0.0)
        x = sv.solution()
        out = {"n": nn, "ranks": world, "ingest_chunks_on_rank0": int(inf.ingest_chunks), "ingest_threads": int(inf.ingest_threads),
               "load_seconds": load_s, "iterations": int(r.iterations)}
        sv.close()
        barrier()
        if rank == 0:
            o = oracle.cg_solve(A, b, k, 0.0)
            out.update(oracle_iterations=int(o.iters), x_rel_l2_vs_oracle=rel_l2(x, o.x), bound=1e-10)
            out["ok"] = bool(out["iterations"] == o.iters and out["x_rel_l2_vs_oracle"] <= 1e-10)
            for p in (pa, pb):
                os.remove(p)
            os.rmdir(tmp)
        return out

    def mixed_storage_case(nn: int, k: int):
        """Option matrix_f32 (opt-in, NOT the headline): the same generate-mode system with the matrix block held as fp32 (0, 1, 2 are
        fp32 numbers, so it is the same problem), vectors / products / sums / scalars fp64; same loop engine and GEMV timing as `value`."""
        sv = make_solver(nn)
        sv.set_option("matrix_f32", 1)
        sv.set_option("loop_mode", 2)
        sv.set_option("time_gemv", 2)
        sv.generate_matrix(nn, nn)
        sv.generate_rhs()
        sv.solve(min(k, 20), 1e-9)
        barrier()
        r = sv.solve(k, 1e-9)
        secs = max_over_ranks(r.solve_seconds)
        ms = 1e3 * max_over_ranks(r.gemv_seconds) / max(r.gemv_launches_timed, 1)
        x = sv.solution()
        inf = sv.info
        rows = int(inf.local_rows)
        out = {"n": nn, "ranks": world, "matrix_elem_bytes": int(inf.matrix_elem_bytes), "matrix_GB_per_gpu": 4.0 * rows * nn / 1e9,
               "entries_rounded": int(inf.matrix_f32_inexact), "gemv_variant": int(inf.gemv_variant),
               "iterations": int(r.iterations), "rel_residual": r.rel_residual, "iterations_per_s": r.iterations_run / secs,
               "gemv_ms": ms, "gemv_GBps_per_gpu": 4.0 * rows * nn / ms / 1e6, "gemv_launches_timed": int(r.gemv_launches_timed),
               "note": "opt-in storage mode, not the headline metric: fp32 matrix block, everything else fp64"}
        sv.close()
        if rank == 0:
            import oracle
            o = oracle.cg_solve_generated(nn, k, 1e-9)
            out.update(oracle_iterations=int(o.iters), oracle_rel_residual=o.rel, x_rel_l2_vs_oracle=rel_l2(x, o.x), bound=1e-12)
            out["ok"] = bool(out["iterations"] == o.iters and out["x_rel_l2_vs_oracle"] <= 1e-12 and out["entries_rounded"] == 0)
        return out

    configs = {}
    mixed = None
    if not args.no_extras:
        if args.gemv_variant in (0, 32, 36):
            mixed = mixed_storage_case(n, iters)
        rem = generate_case(n + 3, 30, 1e-12)
        ing = file_case(1501, 60, 256 << 10, 4)
        if rank == 0:
            parity["remainder_rows"] = rem
            parity["file_ingest"] = ing
        # ---------------- the other BASELINE configs
        if world == 1:
            c0 = generate_case(10000, 1000, 1e-10, time_gemv_reps=50)   # configs[0]; > 1000 iterations: north_star bound (tests/test_gpu_parity.py LONG_RUN)
            c1 = generate_case(50000, 1000, 1e-10, time_gemv_reps=20)   # configs[1]
            c4 = config4_file_mode(lamcg_b200, np, rel_l2)
            configs = {"configs[0] generate n=10000 -i 1000": c0, "configs[1] generate n=50000 -i 1000": c1, "configs[4] file mode n=2048": c4}
        if world == 8:
            c3 = generate_case(300000, 30, 1e-12, time_gemv_reps=10)    # configs[3]: 720 GB, 90 GB per GPU
            if rank == 0:
                configs = {"configs[3] generate n=300000 -i 30 on 8 GPUs": c3}

    # ---------------- traffic of the K1 kernel, measured now (ncu around a probe process on this rank-0 GPU), and the reference GPU class
    traffic = traffic_src = None
    ref_gpu = None
    if rank == 0:
        if not args.no_extras and not args.no_ncu_traffic:
            traffic, traffic_src = measure_k1_traffic(n, world, local_rank)
        if traffic is None:
            tpath = os.path.join(REPO, "profiles", "gemv_traffic.json")
            if os.path.exists(tpath):  # committed ncu capture of this kernel; only valid for the same n and rank count
                with open(tpath) as f:
                    tj = json.load(f)
                if tj.get("n") == n and tj.get("ranks") == world:
                    traffic, traffic_src = tj.get("dram_bytes_per_launch"), "committed capture profiles/gemv_traffic.json (" + (traffic_src or "ncu not run") + ")"
        if world == 1 and not args.no_extras and not args.no_reference_gpu:
            ref_gpu = reference_gpu_block(configs, value)

    if rank == 0:
        peak, peak_src = measured_peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"generate-mode SPD n={n} fp64 ({8.0 * n * n / 1e9:.0f} GB), {iters} CG iterations per step, "
                                   f"row-partitioned over {world} GPU(s) (BASELINE configs[2])",
                       "n": n, "iters_per_step": iters, "rel_error": 1e-9, "ranks": world,
                       "rows_per_gpu": local_rows, "gemv_variant": gemv_variant,
                       "gemv_grid": gemv_grid, "gemv_block": gemv_block, "gemv_smem": gemv_smem,
                       "comm": comm_mode, "comm_note": comm_state["note"],
                       "loop": "value AND e2e: the CUDA-graph loop (the library's default engine at this size) with CUDA events recorded inside the captured "
                               "chunks around one GEMV launch in 16; the same loop without events is in graph_iterations_per_s, plain stream launches "
                               "with events around EVERY GEMV (round 1's timed engine) in stream_loop",
                       "gemv_launches_timed": gemv_timed,
                       "l2": f"inputs larger than L2: {bytes_per_gemv / 1e9:.1f} GB streamed per GEMV per GPU vs 126 MB L2, no flush needed",
                       "generate_seconds": gen_s},
            "gemv_ms": gemv_ms, "gemv_GBps_per_gpu": gemv_gbps, "graph_iterations_per_s": graph_its,
            "wall_s_timed_region": wall_max, "rel_residual_after_step": last.rel_residual,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": gemv_gbps, "peak": peak, "unit": "GB/s", "frac": gemv_gbps / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": f"{K1_KERNEL_REGEX} (K1: Ap = A p + fused p.Ap)",
                         "algorithmic_bytes_per_launch": bytes_per_gemv, "avg_launch_ms": gemv_ms, "launches_timed": gemv_timed,
                         "read_only_stream_GBps": stream_gbps, "frac_of_read_only_stream": gemv_gbps / stream_gbps,
                         "frac_of_nominal_8000": gemv_gbps / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n + 64,
                    "loop": "CUDA-graph loop + sampled GEMV events (same engine as `value`)",
                    "note": "per step: b from pinned host memory (H2D), solve(), x back to pinned host memory (D2H); "
                            "A stays resident in HBM between steps as in the reference's load-once / generate-once flow"},
            "stream_loop": {"iterations_per_s": stream_its, "gemv_ms": stream_gemv_ms,
                            "loop": "plain stream launches with CUDA events around every GEMV (the engine round 1 timed `value` with)"},
            "parity": parity,
            "parity_ok": bool(parity) and all(v.get("ok", False) for v in parity.values()),
        }
        if configs:
            line["configs"] = configs
        if mixed is not None:
            line["matrix_f32"] = mixed
        if ref_gpu is not None:
            line["reference_gpu"] = ref_gpu
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, desc = reference_cpu_block_sample(n, args.ref_iters)
                line["cpu_baseline"] = dict(desc, value=v, unit=UNIT)
            except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def config4_file_mode(lamcg_b200, np, rel_l2):
    """BASELINE configs[4]: file mode, n = 2048, the random_spd_system distribution (seed 42), -i 1000 -e 1e-9.  The system is
    made by the library's GPU generator, written in the reference format, loaded back through lamcg_load_matrix (what the
    reference CLI flow does), solved with the default loop (one cooperative kernel) and compared with the dense CPU oracle
    on the SAME file: stopping iteration, and x at matched iteration count (rel_error = 0)."""
    import oracle
    from oracle import fileformat
    nn = 2048
    tmp = tempfile.mkdtemp(prefix="lamcg_bench_")
    pa, pb = os.path.join(tmp, "A.bin"), os.path.join(tmp, "b.bin")
    sv = lamcg_b200.Solver(0)
    sv.random_spd_system(nn, 42)
    sv.save_system(pa, pb)
    sv.load_matrix(pa)
    sv.load_rhs(pb)
    sv.solve(1000, 1e-9)
    r = sv.solve(1000, 1e-9)
    A, b = fileformat.read_matrix(pa), fileformat.read_vector(pb)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    k = o.iters
    rm = sv.solve(k, 0.0)
    xm = sv.solution()
    om = oracle.cg_solve(A, b, k, 0.0)
    noise = None
    if oracle.ref_available():  # the unmodified reference against itself when only OMP_NUM_THREADS changes (same k iterations)
        xs = [om.x] + [oracle.ref_omp_solve(A, b, k, 0.0, threads=t).x for t in (2, 3, 4, 5, 6, 7, 8, 16)]
        noise = max(rel_l2(xs[i], xs[j]) for i in range(len(xs)) for j in range(i))  # same definition as tests/parity_util.py
    bound = 1e-10 if noise is None or noise < 5e-11 else max(1e-10, 2 * noise)
    # end to end from host buffers: H2D of A (33.5 MB) + b, solve, D2H of x
    import torch
    At, bt = torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory()
    xt = torch.empty(nn, dtype=torch.float64).pin_memory()
    cg = lamcg_b200.ConjugateGradient_B200(0, verbose=False)
    cg.solve_system(At, bt, xt, 1000, 1e-9)
    reps, t0 = 5, time.perf_counter()
    for _ in range(reps):
        cg.solve_system(At, bt, xt, 1000, 1e-9)
    dt = (time.perf_counter() - t0) / reps
    cg.close()
    out = {"n": nn, "iterations": int(r.iterations), "oracle_iterations": int(o.iters),
           "iterations_per_s": r.iterations_run / r.solve_seconds, "us_per_iteration": 1e6 * r.solve_seconds / r.iterations_run,
           "kernel_launches": int(r.kernel_launches), "loop_phase_cycles": sv.loop_profile(),
           "x_rel_l2_vs_oracle_matched_iterations": rel_l2(xm, om.x), "matched_iterations": int(rm.iterations_run),
           "reference_vs_itself_over_threads": noise, "bound": bound,
           "e2e_seconds_per_solve_incl_matrix_upload": dt, "e2e_iterations_per_s": r.iterations_run / dt,
           "h2d_bytes_per_solve": 8 * nn * nn + 8 * nn, "d2h_bytes_per_solve": 8 * nn}
    out["ok"] = bool(abs(out["iterations"] - o.iters) <= 1 and out["x_rel_l2_vs_oracle_matched_iterations"] <= bound)
    sv.close()
    for p in (pa, pb):
        os.remove(p)
    os.rmdir(tmp)
    return out


def measure_k1_traffic(n: int, world: int, device: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the K1 kernel on rank 0's row-block shape, measured now:
    `ncu` wraps tools/traffic_probe.py, which generates the same block (rank 0 of `world`) and launches the kernel 3 times.
    Returns (bytes per launch | None, source string)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{K1_KERNEL_REGEX}",
           "-c", "2", "--csv", sys.executable, os.path.join(REPO, "tools", "traffic_probe.py"), str(n), str(world)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[device] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(device))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    except Exception as e:
        return None, f"ncu failed: {e!r}"
    import csv
    vals = {}
    rows = [ln for ln in res.stdout.splitlines() if ln.startswith('"')]
    for row in csv.DictReader(rows):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "byte").lower()
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
        vals.setdefault(row["ID"], {})[row["Metric Name"]] = v
    per_launch = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in vals.values() if d]
    if not per_launch:
        return None, f"ncu produced no metrics (rc {res.returncode}): {(res.stderr or res.stdout)[-160:]!r}"
    traffic = per_launch[-1]
    try:
        with open(os.path.join(REPO, "profiles", f"gemv_traffic_n{world}.json"), "w") as f:
            json.dump({"n": n, "ranks": world, "dram_bytes_per_launch": traffic, "launches": per_launch,
                       "how": "bench.py: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:" + K1_KERNEL_REGEX + " around tools/traffic_probe.py"}, f, indent=1)
    except OSError:
        pass
    return traffic, "measured in this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum around tools/traffic_probe.py (same kernel, same row block, last of 2 captured launches)"


def reference_gpu_block(configs: dict, our_headline_its: float):
    """The UNMODIFIED reference GPU class (ConjugateGradient_GPU_CUDA<double>, GPU_CUDA.cu:225-316) on this B200, outside the timed
    region and after every handle of ours is closed.  Its solve() mallocs, uploads A from pageable memory, iterates and frees,
    so seconds per iteration = (t(K1) - t(K0)) / (K1 - K0) from two solves on the same resident host matrix."""
    import oracle
    if not oracle.ref_gpu_available("single"):
        return {"unavailable": "oracle/_ref/ref_gpu_single.out not built (needs /root/reference at build time)"}
    tot, avail = host_mem_gb()
    out = {"class": "ConjugateGradient_GPU_CUDA<double> (unmodified, compiled for sm_100 under oracle/_ref)", "host_mem_available_gb": avail, "systems": []}
    sizes = [(50000, 20, 220)]
    if avail >= 96.0:
        sizes.append((100000, 5, 105))
    for nn, k0, k1 in sizes:
        try:
            # every solve() of the reference re-uploads A from pageable memory (2.5 s at n = 50000, 7.5-8 s at n = 100000, +-1 s from
            # run to run): each iteration count is timed twice and the faster run is used
            runs = oracle.ref_gpu_solve("single", [k0, k1, k0, k1], 1e-9, n=nn, timeout=1200)
        except Exception as e:
            out["systems"].append({"n": nn, "failed": repr(e)[-200:]})
            continue
        t0 = [r["seconds"] for r in runs if r["max_iters"] == k0]
        t1 = [r["seconds"] for r in runs if r["max_iters"] == k1]
        if not t0 or not t1:
            out["systems"].append({"n": nn, "failed": "harness printed no timing"})
            continue
        per_it = (min(t1) - min(t0)) / (k1 - k0)
        ours = None
        if nn == 50000:
            c1 = configs.get("configs[1] generate n=50000 -i 1000")
            ours = c1["ms_per_iteration"] if c1 else None
        elif nn == 100000:
            ours = 1e3 / our_headline_its
        row = {"n": nn, "matrix_GB": 8.0 * nn * nn / 1e9, "reference_ms_per_iteration": 1e3 * per_it,
               "reference_effective_GBps": 8.0 * nn * nn / per_it / 1e9, "reference_setup_and_upload_s": min(t0) - k0 * per_it,
               "iterations": [k0, k1], "solve_wall_s": [r["seconds"] for r in runs], "this_library_ms_per_iteration": ours}
        if ours:
            row["loop_speedup"] = 1e3 * per_it / ours
        out["systems"].append(row)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=100, help="CG iterations per step (our arm)")
    ap.add_argument("--ref-iters", type=int, default=20, help="CG iterations per step of the CPU reference row-block sample")
    ap.add_argument("--ref-full-iters", type=int, default=2, help="CG iterations per step of the CPU reference on the whole system")
    ap.add_argument("--ref-block-sample", action="store_true", help="--impl reference: always time the row-block sample, never the whole system")
    ap.add_argument("--gemv-variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip everything after the timed region (parity cases, other configs, ncu traffic, reference GPU)")
    ap.add_argument("--no-ncu-traffic", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--comm", default=os.environ.get("LAMCG_COMM", "peer"), choices=["nccl", "peer"],
                    help="multi-GPU exchange: NCCL collectives or fused NVLink peer stores")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
