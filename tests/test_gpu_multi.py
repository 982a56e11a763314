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
