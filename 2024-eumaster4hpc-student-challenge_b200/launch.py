"""Host-side plumbing for the one-process-per-GPU model (torchrun): rank discovery, the row
partition rule, and the few bytes of bootstrap traffic (NCCL unique id, peer-memory handles) that
the reference moves with MPI_Bcast (GPU/distributed/ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:320-327).
torch.distributed is only the messenger here; nothing on the data path goes through it.

Everything in this file runs on CPU too (gloo), which is how tests/test_host_multirank.py covers it.
"""
from __future__ import annotations

import os


def world_from_env() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) as torchrun exports them; (0, 1, 0) for a plain launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def partition(n: int, nranks: int, rank: int) -> tuple[int, int]:
    """(local_rows, row_offset): n/P rows per rank, the remainder to the last rank — the reference's
    rule (CPU/ConjugateGradient_CPU_MPI_OMP.hpp:175-184), mirrored by lamcg.cu: partition()."""
    base = n // nranks
    return base + (n % nranks if rank == nranks - 1 else 0), base * rank


def broadcast_bytes(payload: bytes | None, src: int = 0, dist=None) -> bytes:
    """Every rank returns rank `src`'s payload."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert payload is not None
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def allgather_bytes(payload: bytes, dist=None) -> list[bytes]:
    """Every rank returns [payload of rank 0, payload of rank 1, ...]."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [payload]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return out


def bootstrap_comm(solver, n: int | None = None, mode: str = "nccl", dist=None, make_id=None) -> str:
    """Give a ranked solver its communicator.  mode "nccl": rank 0 draws the unique id, everyone
    receives it and calls comm_init_nccl.  mode "peer": every rank exports a handle to its exchange
    buffer (needs n), all handles are gathered, every rank imports them.  Returns the mode used."""
    if solver.nranks == 1:
        return "none"
    if mode == "nccl":
        make_id = make_id or type(solver).nccl_unique_id
        uid = broadcast_bytes(make_id() if solver.rank == 0 else None, 0, dist)
        solver.comm_init_nccl(uid)
    elif mode == "peer":
        assert n is not None, "peer mode sizes its exchange buffer from n"
        handles = allgather_bytes(solver.comm_peer_export(n), dist)
        solver.comm_init_peer(b"".join(handles))
    else:
        raise ValueError(f"unknown comm mode {mode!r}")
    return mode


def assemble(slices: list, n: int):
    """Concatenate per-rank solution slices (rank order) and check they tile [0, n)."""
    import numpy as np
    x = np.concatenate(slices)
    assert x.size == n, (x.size, n)
    return x
