"""Multi-GPU parity (needs >= 2 GPUs; skipped on the 1-GPU box): P ranks under torchrun, NCCL and fused
peer-store exchange, against the CPU oracle.  Run with:  gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu"""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("comm", ["nccl", "peer"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranked_solve_matches_oracle(comm, world, tmp_path):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "report.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "mgpu_worker.py"), comm, out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    rep = json.load(open(out))
    assert rep["ok"], rep
    dst = os.path.join(REPO, "gpurun_out")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(dst, f"mgpu_report_{comm}_{world}.json"), "w") as f:
        json.dump(rep, f, indent=1)


@pytest.mark.parametrize("exe_name,fields", [("test_CG_MultiGPUS_CUDA_MPI.out", 9), ("test_CG_MultiGPUS_CUDA_NCCL.out", 10)])
@pytest.mark.parametrize("comm", ["nccl", "peer"])
def test_cpp_driver_forked_ranks(comm, exe_name, fields, tmp_path):
    """The getopt driver with LAMCG_NGPUS=2: RankWorld forks one process per GPU (the reference uses
    srun -n 2), rank 0 prints the CSV line and writes the gathered x."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import math
    import numpy as np
    import oracle
    from oracle import fileformat
    exe = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "test", exe_name)
    sol = str(tmp_path / "sol.bin")
    env = dict(os.environ, LAMCG_NGPUS="2", LAMCG_COMM=comm)
    n, it = 10007, 150
    res = subprocess.run([exe, "-s", str(n), "-i", str(it), "-e", "1e-9", "-o", sol], capture_output=True, text=True, env=env, timeout=200)
    assert res.returncode == 0, res.stdout + res.stderr
    f = res.stdout.strip().split(",")
    shift = fields - 9  # the NCCL-named executable prints the communicator-init seconds after io_s
    assert len(f) == fields and int(f[0]) == n and int(f[1]) == 2 and int(f[6 + shift]) == it + 1
    o = oracle.cg_solve_generated(n, it, 1e-9)
    assert math.isclose(float(f[7 + shift]), o.rel, rel_tol=2e-5)
    x = fileformat.read_vector(sol)
    assert np.linalg.norm(x - o.x) / np.linalg.norm(o.x) <= 1e-12


def test_cpp_positional_multi_gpu_driver(tmp_path):
    """test_CG_MultiGPUS_CUDA.out (positional, file mode) with 2 forked ranks against the oracle."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import oracle
    import parity_util
    from oracle import fileformat, random_spd
    exe = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "test", "test_CG_MultiGPUS_CUDA.out")
    n = 777
    A, b = random_spd.random_spd_system(n, 3)
    pa, pb, px = (str(tmp_path / f) for f in ("A.bin", "b.bin", "x.bin"))
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    for comm in ("peer", "nccl"):
        env = dict(os.environ, LAMCG_NGPUS="2", LAMCG_COMM=comm)
        res = subprocess.run([exe, pa, pb, px, "1000", "1e-9"], capture_output=True, text=True, env=env, timeout=200)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "Converged in" in res.stdout and "Finished successfully" in res.stdout and "GPUs (ranks):      2" in res.stdout
        x = fileformat.read_vector(px)
        assert np.linalg.norm(x - o.x) / np.linalg.norm(o.x) <= 1e-9
        res = subprocess.run([exe, pa, pb, px, str(o.iters), "0"], capture_output=True, text=True, env=env, timeout=200)
        om = oracle.cg_solve(A, b, o.iters, 0.0)
        x = fileformat.read_vector(px)
        assert parity_util.rel_l2(x, om.x) <= parity_util.x_tolerance(parity_util.reference_self_noise(A, b, o.iters, om.x))
