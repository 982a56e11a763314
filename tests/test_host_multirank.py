"""World-size-2 tests of the host-side multi-rank logic on CPU (gloo): the bootstrap that replaces the
reference's MPI_Bcast of the NCCL id, the peer-handle all-gather, the partition rule and the assembly
of per-rank solution slices.  The GPU side is replaced by a recording fake; no CUDA needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


class FakeSolver:
    def __init__(self, rank, nranks):
        self.rank, self.nranks = rank, nranks
        self.got_id = None
        self.got_handles = None

    @staticmethod
    def nccl_unique_id():
        return bytes(range(128))

    def comm_init_nccl(self, uid):
        self.got_id = uid

    def comm_peer_export(self, n):
        return bytes([self.rank]) * 128

    def comm_init_peer(self, blob):
        self.got_handles = blob


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import lamcg_b200
    launch = lamcg_b200.launch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert launch.world_from_env() == (rank, world, rank)
        s = FakeSolver(rank, world)
        ids = []
        mode = launch.bootstrap_comm(s, mode="nccl", dist=dist, make_id=lambda: ids.append(1) or b"\x07" * 128)
        assert mode == "nccl" and s.got_id == b"\x07" * 128
        assert len(ids) == (1 if rank == 0 else 0)  # only rank 0 draws an id
        launch.bootstrap_comm(s, n=n, mode="peer", dist=dist)
        assert s.got_handles == b"".join(bytes([r]) * 128 for r in range(world))
        # every rank solves "its" slice of a known vector; slices must tile [0, n) in rank order
        rows, off = launch.partition(n, world, rank)
        assert (rows, off) == oracle.partition(n, world, rank)
        x_true = np.arange(n, dtype=np.float64)
        slices = launch.allgather_bytes(x_true[off:off + rows].tobytes(), dist)
        x = launch.assemble([np.frombuffer(b, dtype=np.float64) for b in slices], n)
        assert np.array_equal(x, x_true)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 10007])
def test_bootstrap_partition_and_assembly_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_partition_matches_oracle_rule():
    import lamcg_b200
    for n, P in [(10, 3), (100000, 8), (300000, 8), (7, 8), (2048, 4), (10007, 2)]:
        tot = 0
        for r in range(P):
            assert lamcg_b200.launch.partition(n, P, r) == oracle.partition(n, P, r)
            tot += lamcg_b200.launch.partition(n, P, r)[0]
        assert tot == n


def test_single_rank_bootstrap_is_a_noop():
    import lamcg_b200
    s = FakeSolver(0, 1)
    assert lamcg_b200.launch.bootstrap_comm(s) == "none" and s.got_id is None
    assert lamcg_b200.launch.broadcast_bytes(b"x") == b"x"
    assert lamcg_b200.launch.allgather_bytes(b"y") == [b"y"]


def test_set_matrix_shape_validation_happens_before_the_c_call(lamcg):
    """Solver.set_matrix: layout 0 takes the whole (n, n) matrix, layout 1 this rank's (local_rows, n) block; a short or
    non-square array must raise ValueError in Python instead of letting the library's 2-D copy read past the buffer."""
    import types
    check = lamcg.Solver._check_matrix_shape
    single = types.SimpleNamespace(rank=0, nranks=1)
    last_of_3 = types.SimpleNamespace(rank=2, nranks=3)
    assert check(single, (7, 7), 0) == 7
    assert check(last_of_3, (10, 10), 0) == 10
    assert check(last_of_3, (4, 10), 1) == 10          # 10 // 3 = 3 rows + the remainder row on the last rank
    for who, shape, layout in [(single, (6, 7), 0), (single, (7,), 0), (single, (7, 7, 1), 0), (last_of_3, (3, 10), 1),
                               (last_of_3, (10, 10), 1), (single, (0, 0), 0), (single, (7, 7), 2)]:
        with pytest.raises(ValueError):
            check(who, shape, layout)
