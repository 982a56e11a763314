"""lamcg_solve_resume / lamcg_checkpoint_save / lamcg_checkpoint_load (checkpoint of (x, r, p, rr, iter)
for very long generate-mode solves).  The reference has no such call (a second solve() restarts from x = 0,
OMP.hpp:56-67), so the property tested is the strongest one available: an interrupted solve is BIT-IDENTICAL to the
uninterrupted one — same x, same iteration count, same residual history — whether the state stayed on the device or
went through a checkpoint file into a fresh handle.  The uninterrupted solve itself is held to the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import parity_util
from oracle import fileformat, random_spd

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GETOPT = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "test", "test_CG_MultiGPUS_CUDA_MPI.out")


def _generated(lamcg, n, loop_mode, dtype="f64"):
    s = lamcg.Solver(0, dtype=dtype)
    s.set_option("loop_mode", loop_mode)
    s.generate_matrix(n, n)
    s.generate_rhs()
    return s


@pytest.mark.parametrize("loop_mode", [1, 2])
@pytest.mark.parametrize("k,m", [(7, 30), (16, 17), (1, 1), (33, 200)])
def test_resume_is_bit_identical_to_uninterrupted_generate_mode(lamcg, loop_mode, k, m):
    n = 5003
    with _generated(lamcg, n, loop_mode) as s:
        full = s.solve(k + m, 1e-9)
        x_full, h_full = s.solution(), s.residual_history()
        part = s.solve(k, 1e-9)
        assert part.iterations == k + 1 and not part.converged
        res = s.solve_resume(m, 1e-9)
        x_res, h_res = s.solution(), s.residual_history()
    assert (res.iterations, res.iterations_run, res.converged) == (full.iterations, full.iterations_run, full.converged)
    assert res.rel_residual == full.rel_residual
    assert np.array_equal(x_res, x_full)
    assert np.array_equal(h_res, h_full) and len(h_res) == k + m
    o = oracle.cg_solve_generated(n, k + m, 1e-9)
    assert o.iters == res.iterations and parity_util.rel_l2(x_res, o.x) <= 1e-12


def test_chained_resumes_and_convergence_inside_a_resume(lamcg):
    """n = 600 converges at iteration 300 in generate mode: solve(100) + resume(100) + resume(500) stops there too."""
    n = 600
    with _generated(lamcg, n, 2) as s:
        full = s.solve(10000, 1e-9)
        x_full = s.solution()
        assert full.converged and full.iterations == n // 2
        assert not s.solve(100, 1e-9).converged
        assert not s.solve_resume(100, 1e-9).converged
        res = s.solve_resume(500, 1e-9)
        assert res.converged and res.iterations == full.iterations and res.iterations_run == full.iterations_run
        assert np.array_equal(s.solution(), x_full)
        assert len(s.residual_history()) == n // 2
        with pytest.raises(lamcg.LamcgError) as e:   # converged: nothing left to resume
            s.solve_resume(10, 1e-9)
        assert e.value.code == -7


def test_resume_file_mode_spd_matches_oracle(lamcg):
    n = 300
    A, b = random_spd.random_spd_system(n, 9)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    with lamcg.Solver(0) as s:
        s.set_option("loop_mode", 2)
        s.set_matrix(A)
        s.set_rhs(b)
        full = s.solve(1000, 1e-9)
        x_full = s.solution()
        s.solve(101, 1e-9)
        res = s.solve_resume(899, 1e-9)
        assert res.converged and res.iterations == full.iterations
        assert np.array_equal(s.solution(), x_full)
    assert parity_util.iterations_within_one_of_reference(res.iterations, A, b, 1000, 1e-9, o.iters)[0]
    assert parity_util.rel_l2(x_full, o.x) <= 1e-9


@pytest.mark.parametrize("dtype,k", [("f64", 40), ("f64", 41), ("f32", 25)])
def test_checkpoint_roundtrip_into_a_fresh_handle(lamcg, tmp_path, dtype, k):
    n, m = 4100, 60
    ck = str(tmp_path / "cg.ckpt")
    with _generated(lamcg, n, 2, dtype) as s:
        full = s.solve(k + m, 1e-9)
        x_full, h_full = s.solution(), s.residual_history()
        s.solve(k, 1e-9)
        s.checkpoint_save(ck)
    esz = 8 if dtype == "f64" else 4
    assert os.path.getsize(ck) >= 3 * n * esz + 8 * k
    with _generated(lamcg, n, 2, dtype) as t:        # a new process would do exactly this: same system, then the checkpoint
        t.checkpoint_load(ck)
        assert len(t.residual_history()) == k
        res = t.solve_resume(m, 1e-9)
        assert (res.iterations, res.iterations_run) == (full.iterations, full.iterations_run)
        assert res.rel_residual == full.rel_residual
        assert np.array_equal(t.solution(), x_full)
        assert np.array_equal(t.residual_history(), h_full)


def test_resume_and_checkpoint_error_paths(lamcg, tmp_path):
    ck = str(tmp_path / "cg.ckpt")
    with _generated(lamcg, 2000, 2) as s:
        with pytest.raises(lamcg.LamcgError) as e:          # no solve yet
            s.solve_resume(5, 1e-9)
        assert e.value.code == -7
        s.solve(10, 1e-9)
        with pytest.raises(lamcg.LamcgError) as e:
            s.solve_resume(0, 1e-9)
        assert e.value.code == -1
        s.checkpoint_save(ck)
        s.generate_rhs()                                      # the system changed: the kept state is void
        with pytest.raises(lamcg.LamcgError) as e:
            s.solve_resume(5, 1e-9)
        assert e.value.code == -7
        with pytest.raises(lamcg.LamcgError) as e:
            s.checkpoint_save(ck + "2")
        assert e.value.code == -7
        with pytest.raises(lamcg.LamcgError) as e:
            s.checkpoint_load(str(tmp_path / "missing.ckpt"))
        assert e.value.code == -3
        bad = str(tmp_path / "bad.ckpt")
        with open(bad, "wb") as f:
            f.write(b"not a checkpoint at all" * 10)
        with pytest.raises(lamcg.LamcgError) as e:
            s.checkpoint_load(bad)
        assert e.value.code == -3
        with open(ck, "rb") as f:
            blob = f.read()
        with open(bad, "wb") as f:
            f.write(blob[: len(blob) // 2])                   # truncated
        with pytest.raises(lamcg.LamcgError) as e:
            s.checkpoint_load(bad)
        assert e.value.code == -3
    with _generated(lamcg, 2001, 2) as t:                     # another system size
        with pytest.raises(lamcg.LamcgError) as e:
            t.checkpoint_load(ck)
        assert e.value.code == -4
    with _generated(lamcg, 2000, 2, "f32") as t:              # another element type
        with pytest.raises(lamcg.LamcgError) as e:
            t.checkpoint_load(ck)
        assert e.value.code == -4
    with _generated(lamcg, 512, 0) as p:                      # auto -> the one-kernel persistent loop: p never leaves shared memory
        p.solve(10, 1e-9)
        with pytest.raises(lamcg.LamcgError) as e:
            p.solve_resume(5, 1e-9)
        assert e.value.code == -7


def test_cli_checkpoint_and_resume(tmp_path):
    """-c writes the checkpoint when -i runs out; a second process with -r continues; same solution file as one run."""
    n = 5001
    one, two, ck = (str(tmp_path / f) for f in ("one.bin", "two.bin", "cg.ckpt"))
    # the uninterrupted run would pick the one-kernel loop at this size (different summation order): pin it to the graph loop,
    # which -c / -r use, so that the three runs can be compared bit for bit
    r = subprocess.run([GETOPT, "-s", str(n), "-i", "150", "-o", one], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, LAMCG_LOOP_MODE="2"))
    assert r.returncode == 0, r.stderr
    r = subprocess.run([GETOPT, "-s", str(n), "-i", "61", "-o", two, "-c", ck], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and os.path.exists(ck), r.stderr
    r = subprocess.run([GETOPT, "-s", str(n), "-i", "89", "-o", two, "-r", ck], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    f = r.stdout.strip().split(",")
    assert len(f) == 9 and int(f[6]) == 151          # totals: max_iters + 1 of the whole run, like one uninterrupted solve
    assert np.array_equal(fileformat.read_vector(one), fileformat.read_vector(two))
    r = subprocess.run([GETOPT, "-s", str(n + 1), "-i", "5", "-o", two, "-r", ck], capture_output=True, text=True, timeout=300)
    assert r.returncode == 3 and "Failed to read checkpoint" in r.stderr
