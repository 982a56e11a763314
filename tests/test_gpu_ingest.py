"""File-mode ingest pipeline (reference: OMP.hpp:137-197 fread, MPI_OMP.hpp:307-417 MPI-IO with int counts):
T reader threads x 3 pinned staging slots each, chunks pulled off a shared counter, async 2-D H2D copies.  The default chunk is
8 MB, so the small files a test can afford would never leave the single-chunk path; `ingest_chunk_bytes` shrinks the chunk so
that a <= 64 MB file takes the many-chunk, many-thread, slot-reuse path.  Bit equality is checked through the library's own
round trip (load -> save_system == the file) and through the GEMV on integer vectors."""
import os

import numpy as np
import pytest

import oracle
from oracle import fileformat

pytestmark = pytest.mark.gpu


@pytest.fixture()
def solver(lamcg):
    s = lamcg.Solver(0)
    yield s
    s.close()


@pytest.mark.parametrize("n,chunk_bytes,threads", [(2500, 1 << 20, 4),      # 50 MB file, 48 chunks of 52 rows, lda == n
                                                    (2047, 256 << 10, 7),    # lda = 2048 != n: pitched 2-D copies, 128 chunks, odd thread count
                                                    (1000, 1, 16),           # chunk smaller than a row -> one row per chunk, 1000 chunks
                                                    (1500, 1 << 30, 8)])     # chunk larger than the file -> one chunk, one thread
def test_multi_chunk_multi_thread_ingest_is_bit_exact(solver, tmp_path, n, chunk_bytes, threads):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    pa, pb, pa2, pb2 = (str(tmp_path / f) for f in ("A.bin", "b.bin", "A2.bin", "b2.bin"))
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    solver.set_option("ingest_chunk_bytes", chunk_bytes)
    solver.set_option("ingest_threads", threads)
    for _ in range(2):  # the second load reuses the pinned pool and every slot many times over
        solver.load_matrix(pa)
        solver.load_rhs(pb)
        info = solver.info
        rows_per_chunk = max(1, min(chunk_bytes, 256 << 20) // (8 * n)) if chunk_bytes >= 8 * n else 1
        want_chunks = -(-n // min(rows_per_chunk, n))
        assert info.ingest_chunks == want_chunks
        assert info.ingest_threads == min(threads, want_chunks)
        solver.save_system(pa2, pb2)
        assert open(pa, "rb").read() == open(pa2, "rb").read()      # every row landed where it belongs, bit for bit
        assert open(pb, "rb").read() == open(pb2, "rb").read()
    p = (np.arange(n) % 13 - 6).astype(np.float64)
    y, _ = solver.gemv(p)
    bound = 1e-13 * (np.abs(A) @ np.abs(p))
    assert np.all(np.abs(y - oracle.gemv(A, p)) <= bound)


def test_row_block_ingest_of_a_ranked_handle(lamcg, tmp_path):
    """Each rank reads only its own row block at its own offset (MPI_OMP.hpp:376-406), remainder rows on the last rank."""
    n, P = 1003, 4
    rng = np.random.default_rng(5)
    A = rng.integers(-9, 10, size=(n, n)).astype(np.float64)
    pa = str(tmp_path / "A.bin")
    fileformat.write_matrix(pa, A)
    p = rng.integers(-9, 10, size=n).astype(np.float64)
    want = oracle.gemv(A, p)
    for rank in range(P):
        rows, off = oracle.partition(n, P, rank)
        s = lamcg.Solver(0, rank, P)
        s.set_option("ingest_chunk_bytes", 64 << 10)
        s.set_option("ingest_threads", 3)
        s.load_matrix(pa)
        assert (s.info.local_rows, s.info.row_offset) == (rows, off)
        y, _ = s.gemv(p)
        assert np.array_equal(y, want[off:off + rows])
        s.close()


def test_offsets_beyond_2_to_31_elements(lamcg, tmp_path):
    """The reference's MPI-IO loader counts elements in an int: a row block that starts beyond 2^31 elements (n = 50000 on one
    rank) silently reads garbage and the run ends in `10001,-nan` (TESTS/BEST_RESULTS:114).  Here every offset is 64-bit: a
    SPARSE n = 50000 file (20 GB logical, only the last 100 rows and the header are ever written, ~40 MB on disk) is loaded by
    the last of 500 ranks, whose block starts at element 2 495 000 000 > 2^31 (byte offset ~19.96 GB > 2^34)."""
    n, P = 50000, 500
    rank = P - 1
    rows, off = oracle.partition(n, P, rank)
    assert off * n > 2 ** 31 and rows == 100
    rng = np.random.default_rng(11)
    block = rng.integers(-4, 5, size=(rows, n)).astype(np.float64)
    pa = str(tmp_path / "A.bin")
    with open(pa, "wb") as f:
        np.array([n, n], dtype=np.uint64).tofile(f)
        f.truncate(16 + 8 * n * n)
        f.seek(16 + 8 * off * n)
        block.tofile(f)
    if os.stat(pa).st_blocks * 512 > (1 << 30):
        pytest.skip("file system does not keep the file sparse")
    s = lamcg.Solver(0, rank, P)
    s.set_option("ingest_chunk_bytes", 4 << 20)
    s.set_option("ingest_threads", 4)
    s.load_matrix(pa)
    assert (s.info.local_rows, s.info.row_offset) == (rows, off) and s.info.ingest_chunks == 10
    p = rng.integers(-4, 5, size=n).astype(np.float64)
    y, _ = s.gemv(p)
    assert np.array_equal(y, block @ p)                      # small integers: exact in any summation order
    s.close()
    # and a rank whose block lies in the hole reads zeros, not an error
    s = lamcg.Solver(0, 250, P)
    s.load_matrix(pa)
    y, _ = s.gemv(p)
    assert not y.any()
    s.close()
