"""fp32 instantiation: the reference instantiates its GPU classes for float as well
(GPU/local/ConjugateGradient_MultiGPUS_CUDA.cu:539, distributed/*.cu:707,767).  Here float is the STORAGE type
(A, b, x, work vectors, files, caller buffers) while every reduction, alpha, beta and the stop test stay in
fp64 — so the result is at least as accurate as the reference's all-float loop.  Tolerances are fp32-sized:
integer inputs bit-exact; x within 1e-4 of the all-float oracle (which itself sits ~2e-5 from the fp64 oracle)."""
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import random_spd

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture()
def solver32(lamcg):
    s = lamcg.Solver(0, dtype="f32")
    yield s
    s.close()


@pytest.mark.parametrize("variant", [0, 32, 36, 42, 46])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 33, 255, 257, 1000, 1025, 4099])
def test_gemv_f32_integer_inputs_bit_exact(solver32, variant, n):
    rng = np.random.default_rng(7 * n + variant)
    A = rng.integers(-8, 9, size=(n, n)).astype(np.float32)
    p = rng.integers(-8, 9, size=n).astype(np.float32)
    solver32.set_option("gemv_variant", variant)
    solver32.set_matrix(A)
    y, d = solver32.gemv(p)
    y_ref = (A.astype(np.float64) @ p.astype(np.float64))
    assert y.dtype == np.float32 and np.array_equal(y.astype(np.float64), y_ref)
    assert d == float(p.astype(np.float64) @ y_ref)


def test_f32_variant_restrictions(solver32, lamcg):
    solver32.generate_matrix(64, 64)
    with pytest.raises(lamcg.LamcgError):
        solver32.set_option("gemv_variant", 11)      # the ldg / TMA families are fp64 only
    solver32.set_option("gemv_variant", 0)
    solver32.generate_rhs()
    solver32.set_option("loop_mode", 3)              # so is the persistent loop
    with pytest.raises(lamcg.LamcgError):
        solver32.solve(10, 1e-4)
    assert solver32.info.dtype == 1


@pytest.mark.parametrize("n,max_iters", [(8, 100), (1000, 100), (2048, 300), (10007, 150)])
@pytest.mark.parametrize("loop_mode", [1, 2])
def test_generate_mode_f32(solver32, n, max_iters, loop_mode):
    solver32.set_option("loop_mode", loop_mode)
    solver32.generate_matrix(n, n)
    solver32.generate_rhs()
    r = solver32.solve(max_iters, 1e-5)
    x = solver32.solution()
    assert x.dtype == np.float32
    o32 = oracle.cg_solve_f32(None, np.ones(n), max_iters, 1e-5)
    o64 = oracle.cg_solve_generated(n, max_iters, 1e-5)
    assert abs(r.iterations - o64.iters) <= 1, (r.iterations, o64.iters, o32.iters)
    assert rel_l2(x, o64.x) <= 1e-4 and rel_l2(x, o32.x) <= 2e-4
    if not o64.converged:  # a finite-termination step bottoms out at the fp32 floor (~1e-7), not at 1e-17
        assert math.isclose(r.rel_residual, o64.rel, rel_tol=1e-2)


def test_file_mode_f32_matches_reference_float_class(solver32, tmp_path):
    """float files (header + float32 data, the reference's sizeof(FloatingType) layout) -> load -> solve -> save."""
    n = 300
    A, b = random_spd.random_spd_system(n, 9)
    A32, b32 = A.astype(np.float32), b.astype(np.float32)
    pa, pb, px = (str(tmp_path / f) for f in ("A.bin", "b.bin", "x.bin"))
    for path, arr, shape in ((pa, A32, (n, n)), (pb, b32, (n, 1))):
        with open(path, "wb") as f:
            np.array(shape, dtype=np.uint64).tofile(f)
            arr.tofile(f)
    solver32.load_matrix(pa)
    solver32.load_rhs(pb)
    r = solver32.solve(1000, 1e-4)
    x = solver32.solution()
    o32 = oracle.cg_solve_f32(A32, b32, 1000, 1e-4)
    if oracle.ref_available():  # the unmodified reference, ConjugateGradient_CPU_OMP<float>
        ref = oracle.ref_omp_solve_f32(A32, b32, 1000, 1e-4, threads=1)
        assert ref.iters == o32.iters and np.array_equal(ref.x, o32.x)
    assert r.converged and abs(r.iterations - o32.iters) <= max(5, o32.iters // 10)
    x_true = np.linalg.solve(A32.astype(np.float64), b32.astype(np.float64))
    assert rel_l2(x, x_true) <= 2e-3 and rel_l2(x, x_true) <= 1.5 * rel_l2(o32.x, x_true) + 1e-4
    solver32.save_solution(px)
    assert os.path.getsize(px) == 16 + 4 * n
    hdr = np.fromfile(px, dtype=np.uint64, count=2)
    assert tuple(hdr) == (n, 1)
    assert np.array_equal(np.fromfile(px, dtype=np.float32, offset=16), x)


def test_cpp_float_instantiation():
    """LAM::ConjugateGradient_B200<float> compiles and solves (the reference instantiates <float> too)."""
    exe = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "test", "test_float_instantiation.out")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "float ok" in res.stdout and "double ok" in res.stdout
