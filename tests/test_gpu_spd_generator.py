"""GPU random SPD generator against the numpy restatement of the reference's
generator (oracle/random_spd.py): same glibc streams, same distribution; the matrices agree to
rounding because recursive Gram-Schmidt and sign-fixed QR produce the same Q for a well-conditioned
random matrix."""
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
import parity_util
from oracle import fileformat, random_spd

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# 1..33: a single CholeskyQR2 panel / the first recursion step around the 32-column leaf; 64, 65, 96: two levels; 257, 1000: the
# fp64 tensor-core GEMM (results of 64 x 64 and more) with ragged tiles and split-K
@pytest.mark.parametrize("n,seed", [(1, 3), (2, 3), (7, 5), (31, 2), (32, 4), (33, 6), (64, 8), (65, 9), (96, 7), (257, 11), (1000, 42)])
def test_generator_matches_numpy_restatement(lamcg, tmp_path, n, seed):
    s = lamcg.Solver(0)
    s.random_spd_system(n, seed)
    pa, pb = str(tmp_path / "A.bin"), str(tmp_path / "b.bin")
    s.save_system(pa, pb)
    A = fileformat.read_matrix(pa)
    b = fileformat.read_vector(pb)
    assert os.path.getsize(pa) == 16 + 8 * n * n and fileformat.read_header(pb) == (n, 1)
    A_ref, b_ref = random_spd.random_spd_system(n, seed)
    assert np.array_equal(b, b_ref)                      # same glibc stream, bit for bit
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
    assert np.linalg.norm(A - A_ref) <= 1e-9 * np.linalg.norm(A_ref)
    w = np.linalg.eigvalsh((A + A.T) / 2)
    assert w.min() >= math.exp(-3.5) * (1 - 1e-6) and w.max() <= math.exp(3.5) * (1 + 1e-6)
    # the generated system is immediately solvable on the same handle, like the reference flow generate -> solve
    r = s.solve(1000, 1e-9)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    assert r.converged and parity_util.iterations_within_one_of_reference(r.iterations, A, b, 1000, 1e-9, o.iters)[0]
    assert parity_util.rel_l2(s.solution(), o.x) <= 1e-9
    s.close()


@pytest.mark.parametrize("n", [200, 777])
def test_tensor_core_and_simt_generators_agree(lamcg, tmp_path, n):
    """The default generator (DMMA products, CholeskyQR2 leaves, device-side glibc stream) against the round-1 algorithm
    (option spd_simt: SIMT products, recursion to single columns): the same matrix to 1e-12 (both are fp64; only summation
    orders and the leaf algorithm differ, and the thin QR factor with positive diagonal is unique), bit-identical rhs.  The oracle
    for both stays the numpy restatement (parity unpinned: the reference needs Intel MKL)."""
    mats = {}
    for simt in (0, 1):
        s = lamcg.Solver(0)
        s.set_option("spd_simt", simt)
        s.random_spd_system(n, 13)
        pa, pb = str(tmp_path / f"A{simt}.bin"), str(tmp_path / f"b{simt}.bin")
        s.save_system(pa, pb)
        mats[simt] = (fileformat.read_matrix(pa), fileformat.read_vector(pb))
        s.close()
    assert np.array_equal(mats[0][1], mats[1][1])
    assert np.linalg.norm(mats[0][0] - mats[1][0]) <= 1e-12 * np.linalg.norm(mats[1][0])


def test_generator_cli_and_reference_cli_reads_it(tmp_path):
    exe = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "random_spd_system.out")
    pa, pb, px = (str(tmp_path / f) for f in ("A.bin", "b.bin", "x.bin"))
    res = subprocess.run([exe, "300", pa, pb, "42"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "Finished successfully" in res.stdout, res.stdout + res.stderr
    A, b = fileformat.read_matrix(pa), fileformat.read_vector(pb)
    A_ref, b_ref = random_spd.random_spd_system(300, 42)
    assert np.array_equal(b, b_ref) and np.linalg.norm(A - A_ref) <= 1e-9 * np.linalg.norm(A_ref)
    if os.path.exists(oracle.REF_TEST_OMP):  # the unmodified reference solver accepts the generated files
        r = subprocess.run([oracle.REF_TEST_OMP, pa, pb, px, "1000", "1e-9"], capture_output=True, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=120)
        assert r.returncode == 0 and "Converged in" in r.stdout
    res = subprocess.run([exe, "0", pa, pb, "1"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 1 and "Wrong argument value" in res.stderr
