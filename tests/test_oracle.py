"""CPU tests of the oracle itself: it must agree with the reference's golden vectors before it is
allowed to judge the CUDA path.  No GPU needed."""
import math
import os

import numpy as np
import pytest

import oracle
from oracle import fileformat, random_spd


def test_generate_mode_golden_iterations_and_residual(golden):
    """Reference class output (harness, unmodified sources) vs our restatement: iteration count exact,
    residual to the 6 digits the reference prints, x bit-identical at generation time."""
    for g in golden["generate_mode"]:
        assert g["oracle_bit_identical_x"], g
        o = oracle.cg_solve_generated(g["n"], g["max_iters"], g["rel_error"])
        assert o.iters == g["iters"], g
        assert o.converged == g["converged"], g
        if g["rel_printed"] > 0:
            assert math.isclose(o.rel, g["rel_printed"], rel_tol=2e-5), g
        assert math.isclose(float(np.linalg.norm(o.x)), g["x_norm2"], rel_tol=1e-15), g
        assert math.isclose(float(o.x.sum()), g["x_sum"], rel_tol=1e-15), g


@pytest.mark.parametrize("n", [8, 1000, 2048])
def test_generate_mode_golden_x_bitwise(n, golden_dir):
    x_ref = np.load(os.path.join(golden_dir, f"gen_x_n{n}.npy"))
    max_iters = 100 if n == 8 else 10000
    o = oracle.cg_solve_generated(n, max_iters, 1e-9)
    assert np.array_equal(o.x, x_ref)


def test_generate_mode_cli_golden(golden):
    for g in golden["generate_mode_cli"]:
        assert g["csv_fields"] == 9 and g["n_field"] == g["n"] and g["ranks"] == 1
        o = oracle.cg_solve_generated(g["n"], g["max_iters"], 1e-9)
        assert o.iters == g["iters"]
        assert math.isclose(o.rel, g["rel_printed"], rel_tol=2e-5)


def test_reference_result_dump_known_answers(golden):
    """Rows of the reference's own result files (other hardware, same arithmetic)."""
    for g in golden["reference_result_dumps"]:
        o = oracle.cg_solve_generated(g["n"], g["max_iters"], 1e-9)
        assert o.iters == g["iters"], g
        assert math.isclose(o.rel, g["rel_printed"], rel_tol=2e-5), (g, o.rel)


def test_closed_form_residual():
    """rel after k iterations ~ 1/(k*sqrt(8n)) for k << n/2."""
    n = 20000
    o = oracle.cg_solve_generated(n, 50, 1e-9, history=True)
    for k in (1, 5, 15, 50):
        assert math.isclose(o.hist[k - 1], 1.0 / (k * math.sqrt(8 * n)), rel_tol=2e-3)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 33, 1000, 1025])
def test_structured_gemv_is_bitwise_dense_gemv(n):
    rng = np.random.default_rng(n)
    p = rng.standard_normal(n)
    A = oracle.generate_matrix(n)
    assert np.array_equal(oracle.gemv(A, p), oracle.gemv_generated(p))


@pytest.mark.parametrize("n", [1, 2, 3, 8, 257, 1000])
def test_structured_solve_is_bitwise_dense_solve(n):
    A, b = oracle.generate_matrix(n), oracle.generate_rhs(n)
    d = oracle.cg_solve(A, b, 10000, 1e-9, history=True)
    s = oracle.cg_solve_generated(n, 10000, 1e-9, history=True)
    assert d.iters == s.iters and d.rel == s.rel
    assert np.array_equal(d.x, s.x) and np.array_equal(d.hist, s.hist)
    assert d.iters == (n + 1) // 2  # b symmetric => terminates at ceil(n/2)


def test_n8_fixture():
    o = oracle.cg_solve_generated(8, 100, 1e-9)
    assert o.iters == 4
    np.testing.assert_allclose(o.x * 9, [4, 1, 3, 2, 2, 3, 1, 4], rtol=1e-14)


def test_generator_matches_definition():
    n = 37
    A = oracle.generate_matrix(n)
    E = 2 * np.eye(n) + np.eye(n, k=1) + np.eye(n, k=-1)
    assert np.array_equal(A, E)
    rows, off = oracle.partition(n, 4, 2)
    assert np.array_equal(oracle.generate_matrix(n, rows, off), E[off:off + rows])


@pytest.mark.parametrize("n,P", [(10, 1), (10, 3), (100000, 8), (7, 8), (300000, 8), (2048, 4)])
def test_partition_rule(n, P):
    """n/P rows each, remainder to the last rank (MPI_OMP.hpp:175-184)."""
    covered = 0
    for r in range(P):
        rows, off = oracle.partition(n, P, r)
        assert off == r * (n // P)
        assert rows == n // P + (n % P if r == P - 1 else 0)
        covered += rows
    assert covered == n


def test_file_mode_golden(golden, golden_dir):
    for g in golden["file_mode"]:
        n = g["n"]
        A = fileformat.read_matrix(os.path.join(golden_dir, f"spd_n{n}_A.bin"))
        b = fileformat.read_vector(os.path.join(golden_dir, f"spd_n{n}_b.bin"))
        x_ref = fileformat.read_vector(os.path.join(golden_dir, f"spd_n{n}_x.bin"))
        assert A.shape == (n, n) and b.shape == (n,)
        o = oracle.cg_solve(A, b, 1000, 1e-9)
        assert o.iters == g["iters"] and o.converged
        assert math.isclose(o.rel, g["rel_printed"], rel_tol=2e-6)
        assert np.array_equal(o.x, x_ref)  # reference CLI at OMP_NUM_THREADS=1 == sequential restatement


def test_fileformat_roundtrip_and_garbage_header(tmp_path):
    rng = np.random.default_rng(0)
    M = rng.standard_normal((5, 5))
    p = str(tmp_path / "m.bin")
    fileformat.write_matrix(p, M)
    assert os.path.getsize(p) == 16 + 8 * 25
    assert np.array_equal(fileformat.read_matrix(p), M)
    # a solution file as the reference writes it: int 1 stored with sizeof(size_t) => garbage upper bits
    v = rng.standard_normal(5)
    q = str(tmp_path / "x.bin")
    with open(q, "wb") as f:
        np.array([5, 0x7FFF00000001], dtype=np.uint64).tofile(f)
        v.tofile(f)
    assert np.array_equal(fileformat.read_vector(q), v)


def test_random_spd_distribution():
    n = 96
    A, b = random_spd.random_spd_system(n, 7)
    assert np.allclose(A, A.T, rtol=0, atol=1e-12)
    w = np.linalg.eigvalsh(A)
    assert w.min() >= math.exp(-3.5) * (1 - 1e-9) and w.max() <= math.exp(3.5) * (1 + 1e-9)
    assert np.all(np.abs(b) <= 1.0)
    # same glibc stream as random_spd_system.cpp: srand(seed); 2*rand()/RAND_MAX - 1
    assert np.array_equal(random_spd.random_matrix(n, 1, 17).reshape(-1), oracle.rand_fill(n, 17))
    A2, b2 = random_spd.random_spd_system(n, 7)
    assert np.array_equal(A, A2) and np.array_equal(b, b2)


def test_primitives_definitions():
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(1001), rng.standard_normal(1001)
    seq = 0.0
    for a, c in zip(x, y):
        seq += a * c
    assert oracle.dot(x, y) == seq
    y2 = y.copy()
    oracle.axpby(0.3, x, -1.7, y2)
    assert np.array_equal(y2, 0.3 * x + (-1.7) * y)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_agrees_with_oracle():
    """When the compiled reference is present, re-check the pin live (1 thread: deterministic)."""
    for n, it in [(8, 100), (777, 10000), (2048, 15)]:
        r = oracle.ref_gen_solve(n, it, 1e-9, threads=1)
        o = oracle.cg_solve_generated(n, it, 1e-9)
        assert r.iters == o.iters and np.array_equal(r.x, o.x)
    A, b = random_spd.random_spd_system(128, 3)
    r = oracle.ref_omp_solve(A, b, 1000, 1e-9, threads=1)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    assert r.iters == o.iters and np.array_equal(r.x, o.x)


def test_reference_gpu_harness_is_built_and_refuses_to_run_without_a_gpu():
    """oracle/_ref/ref_gpu_{single,multi}.out (the unmodified reference GPU classes behind oracle/ref_gpu_harness.cu) exist where
    /root/reference was available at build time, print their usage, and stop with exit code 3 when there is no CUDA device —
    the checker has no CPU path either."""
    import subprocess
    import torch
    for variant in ("single", "multi"):
        exe = oracle.REF_GPU_HARNESS[variant]
        if not os.path.exists(exe):
            pytest.skip("oracle/_ref GPU harness not built (needs /root/reference)")
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert r.returncode == 64 and "usage:" in r.stderr
        if not torch.cuda.is_available():
            r = subprocess.run([exe, "gen", "8", "1e-9", "-", "1"], capture_output=True, text=True, timeout=60)
            assert r.returncode == 3 and "no CUDA device" in r.stderr
