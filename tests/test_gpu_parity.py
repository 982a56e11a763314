"""Parity tests proper: the CUDA path (through the C ABI, ctypes -> liblamcg.so) against the CPU
oracle and the reference's golden vectors.  Run on the B200 box:  pytest -m gpu

Tolerances (BASELINE.json north_star): same iteration count +-1 to reach rel_err 1e-9 and solution
relative L2 difference <= 1e-10 in fp64.  Generate mode is the sharp case (the reference's own
thread-count noise there is ~1e-14), so it is held to exact iteration counts, residual to the 6
digits the reference prints, and x to 1e-12.  Integer-valued inputs are held to bit equality.
"""
import json
import math
import os

import numpy as np
import pytest

import oracle
import parity_util
from oracle import fileformat, random_spd

pytestmark = pytest.mark.gpu

X_TOL = 1e-10        # north_star
X_TOL_GEN = 1e-12    # generate mode, see module docstring
LONG_RUN = 1000      # beyond this many iterations rounding differences have been amplified through
                     # thousands of recurrences (the reference's own residual at n/2 termination is only
                     # ~1e-10 for n = 5001), so the north_star tolerance applies instead of the sharp one
REL_TOL = 2e-6       # the reference prints 6 significant digits
X_TOL_STOPPED = 1e-9 # x of two runs that stopped on DIFFERENT iterations: near rel_err 1e-9 one CG step moves x by
                     # ~1e-10 relative, and the unmodified reference differs from itself by up to 2.1e-10 on the
                     # n = 300 fixture of test_gpu_cli.py when only OMP_NUM_THREADS changes (258 vs 262 iterations);
                     # the north_star 1e-10 is therefore asserted at MATCHED iteration count (rel_error = 0)

REPORT = {}


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def check_matched_iterations(solver, A, b, k, tag):
    """x parity at matched iteration count: both sides run exactly k iterations (rel_error = 0 never stops early).  The bound is
    north_star's 1e-10 unless the unmodified reference, run right here with other OMP_NUM_THREADS, is itself noisier than 5e-11
    (then 2 x its own spread): tests/parity_util.py."""
    r = solver.solve(k, 0.0)
    o = oracle.cg_solve(A, b, k, 0.0)
    assert r.iterations == o.iters == k + 1 and r.iterations_run == k
    x = solver.solution()
    err = rel_l2(x, o.x)
    noise = parity_util.reference_self_noise(A, b, k, o.x)
    tol = parity_util.x_tolerance(noise)
    ours_true, oracle_true = parity_util.as_accurate_as_reference(A, b, x, o.x)
    REPORT[f"{tag}_matched_{k}_its"] = {"x_ours_vs_oracle": err, "asserted_bound": tol, "meets_1e-10": bool(err <= X_TOL),
                                        "reference_vs_itself_over_threads": noise,
                                        "ours_vs_exact": ours_true, "oracle_vs_exact": oracle_true}
    assert err <= tol, (err, tol, noise)
    assert ours_true <= 2.0 * oracle_true + 1e-12, (ours_true, oracle_true)
    assert o.rel / 3 <= r.rel_residual <= o.rel * 3  # near rel_err 1e-9 the residual itself wobbles by tens of %


def check_stopping_iteration(r, A, b, o, tag):
    """Iteration count to reach rel_err 1e-9: within +-1 of the unmodified reference's own envelope over OMP_NUM_THREADS, measured now."""
    ok, env = parity_util.iterations_within_one_of_reference(r.iterations, A, b, 1000, 1e-9, o.iters)
    REPORT[f"{tag}_iters"] = {"ours": r.iterations, "oracle": o.iters, "reference_envelope_over_threads": list(env)}
    assert r.converged and ok, (r.iterations, o.iters, env)


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


@pytest.fixture()
def solver(lamcg):
    s = lamcg.Solver(0)
    yield s
    s.close()


# ------------------------------------------------------------------------------------- K1: GEMV
# the six K1 shapes the library keeps (lamcg.cu: make_plan): row sweep 128-bit (32, 36 = defaults) and 256-bit (42, 46) loads,
# warp-per-rows with TMA-staged p (11), TMA ring (2)
VARIANTS = [32, 36, 42, 46, 11, 2]


@pytest.mark.parametrize("variant", VARIANTS)
# 1..9 and 15..17: every tail decomposition (4 + 2 + 1 rows) of the row sweep in a single CTA; 1000 / 1184 / 1185 / 4099: 6-8 and 27-28
# rows per CTA on 148 SMs (full passes + every tail); 255..257, 1025, 2048: chunk and padding edges
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 33, 255, 256, 257, 1000, 1025, 1184, 1185, 2048, 4099])
def test_gemv_integer_inputs_bit_exact(solver, variant, n):
    """Integer-valued A and p: every partial sum is exact, so any summation order must reproduce the
    oracle's sequential sum bit for bit.  Catches indexing, padding and tail bugs at ragged sizes."""
    rng = np.random.default_rng(1000 * variant + n)
    A = rng.integers(-8, 9, size=(n, n)).astype(np.float64)
    p = rng.integers(-8, 9, size=n).astype(np.float64)
    solver.set_option("gemv_variant", variant)
    solver.set_matrix(A)
    y, d = solver.gemv(p)
    y_ref = oracle.gemv(A, p)
    assert np.array_equal(y, y_ref)
    assert d == oracle.dot(p, y_ref)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [257, 1500, 4096])
def test_gemv_random_inputs(solver, variant, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    p = rng.standard_normal(n)
    solver.set_option("gemv_variant", variant)
    solver.set_matrix(A)
    y, d = solver.gemv(p)
    y_ref = oracle.gemv(A, p)
    bound = 1e-13 * (np.abs(A) @ np.abs(p))  # different summation order only
    assert np.all(np.abs(y - y_ref) <= bound)
    assert abs(d - oracle.dot(p, y_ref)) <= 1e-12 * float(np.abs(p) @ np.abs(y_ref))


@pytest.mark.parametrize("variant", VARIANTS)
def test_gemv_generated_matrix_matches_generator(solver, variant):
    """Device generator (MPI_OMP.hpp:237-247) + GEMV on an integer vector == oracle, bit exact."""
    n = 3001
    solver.set_option("gemv_variant", variant)
    solver.generate_matrix(n, n)
    p = np.arange(n, dtype=np.float64) % 17 - 8
    y, _ = solver.gemv(p)
    assert np.array_equal(y, oracle.gemv_generated(p))
    assert np.array_equal(y, oracle.gemv(oracle.generate_matrix(n), p))


# ------------------------------------------------------------------------- K2 / K3: vector kernels
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("n", [1, 2, 31, 256, 257, 1000, 4099, 151553])
def test_vector_kernels_in_isolation(solver, n, fused):
    """K2 (x += alpha p, r -= alpha Ap, r.r) and K3 (beta, p = r + beta p) on their own, as one cooperative launch (fused) and as
    the two kernels NCCL mode uses, against the oracle's axpby / dot (OMP.hpp:233-244, 219-231) on random vectors: the elementwise
    updates are the same unfused operations in the same order, so x, r and p must agree BIT FOR BIT given the same scalars; r.r
    is a sum in another order (1e-14 relative).  n = 151553 = 592 CTAs x 256 threads + 1: every CTA of the largest grid and a
    grid-stride tail."""
    rng = np.random.default_rng(n)
    x, r, p, Ap = (rng.standard_normal(n) for _ in range(4))
    rr = oracle.dot(r, r)
    pAp = abs(oracle.dot(p, Ap)) + 1.0
    xg, rg, pg, alpha, rr_new, beta = solver.vector_update_step(x, r, p, Ap, rr, pAp, fused)
    assert alpha == rr / pAp
    xo, ro, po = x.copy(), r.copy(), p.copy()
    oracle.axpby(alpha, p, 1.0, xo)        # x = alpha p + x
    oracle.axpby(-alpha, Ap, 1.0, ro)      # r = -alpha Ap + r
    assert np.array_equal(xg, xo) and np.array_equal(rg, ro)
    assert abs(rr_new - oracle.dot(ro, ro)) <= 1e-14 * oracle.dot(ro, ro)
    assert beta == rr_new / rr
    oracle.axpby(1.0, ro, beta, po)        # p = r + beta p
    assert np.array_equal(pg, po)


# ------------------------------------------------------------------------------ generate mode
def test_generate_mode_golden(solver, golden):
    """Every generate-mode row of tests/golden/golden.json (produced by the unmodified reference)."""
    for g in golden["generate_mode"]:
        n = g["n"]
        solver.generate_matrix(n, n)
        solver.generate_rhs()
        r = solver.solve(g["max_iters"], g["rel_error"])
        assert r.iterations == g["iters"], (g, r.iterations)
        assert bool(r.converged) == g["converged"], g
        x = solver.solution()
        tol = X_TOL_GEN if g["iters"] <= LONG_RUN else X_TOL
        assert math.isclose(float(np.linalg.norm(x)), g["x_norm2"], rel_tol=tol), g
        if not g["converged"]:
            assert math.isclose(r.rel_residual, g["oracle_rel"], rel_tol=REL_TOL), (g, r.rel_residual)
        else:
            assert r.rel_residual < g["rel_error"]


@pytest.mark.parametrize("n", [8, 1000, 2048])
def test_generate_mode_golden_x(solver, golden_dir, n):
    x_ref = np.load(os.path.join(golden_dir, f"gen_x_n{n}.npy"))  # the reference's private _x
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    r = solver.solve(100 if n == 8 else 10000, 1e-9)
    assert r.converged and r.iterations == (n + 1) // 2
    err = rel_l2(solver.solution(), x_ref)
    REPORT[f"gen_x_rel_l2_n{n}"] = err
    assert err <= (X_TOL_GEN if r.iterations <= LONG_RUN else X_TOL)


@pytest.mark.parametrize("n,max_iters", [(1, 5), (2, 5), (3, 5), (7, 50), (1025, 100), (5001, 10000), (10007, 200), (10000, 1000)])
def test_generate_mode_vs_oracle(solver, n, max_iters):
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    r = solver.solve(max_iters, 1e-9)
    o = oracle.cg_solve_generated(n, max_iters, 1e-9, history=True)
    assert r.iterations == o.iters and bool(r.converged) == o.converged
    x = solver.solution()
    err = rel_l2(x, o.x)
    REPORT[f"gen_vs_oracle_x_rel_l2_n{n}_i{max_iters}"] = err
    assert err <= (X_TOL_GEN if o.iters <= LONG_RUN else X_TOL)
    h = solver.residual_history()
    assert len(h) == min(o.iters, max_iters) == r.iterations_run
    # the last step of an n/2 "finite termination" run is pure rounding noise (1e-6 -> 1e-10 in one step)
    big = o.hist > 1e-9
    np.testing.assert_allclose(h[big], o.hist[big], rtol=REL_TOL)


def test_generate_mode_config1_reference_csv_row(solver, golden):
    """BASELINE config 1: -s 10000 -i 1000 -e 1e-9  ->  the reference prints 1001, 3.53553e-06."""
    g = [e for e in golden["generate_mode_cli"] if e["n"] == 10000 and e["max_iters"] == 1000][0]
    solver.generate_matrix(10000, 10000)
    solver.generate_rhs()
    r = solver.solve(1000, 1e-9)
    assert r.iterations == g["iters"] == 1001 and not r.converged
    assert math.isclose(r.rel_residual, g["rel_printed"], rel_tol=2e-5)


@pytest.mark.parametrize("n,max_iters", [(50000, 15), (100000, 15), (100000, 200)])
def test_full_size_configs_vs_structured_oracle(solver, golden, n, max_iters):
    """BASELINE configs 2 and 3 at full size (20 GB / 80 GB on one B200).  The oracle follows in O(n)
    memory (cg_oracle.c: matvec_generated, bit-identical to the dense loop); known answers from the
    reference's own dumps (TESTS/BEST_RESULTS:184: 100000 -> 16, 7.45356e-05)."""
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    r = solver.solve(max_iters, 1e-9)
    o = oracle.cg_solve_generated(n, max_iters, 1e-9, history=True)
    assert r.iterations == o.iters == max_iters + 1
    assert math.isclose(r.rel_residual, o.rel, rel_tol=REL_TOL)
    if n == 100000 and max_iters == 15:
        assert math.isclose(r.rel_residual, 7.45356e-05, rel_tol=2e-5)
    err = rel_l2(solver.solution(), o.x)
    REPORT[f"full_size_x_rel_l2_n{n}_i{max_iters}"] = err
    assert err <= X_TOL_GEN
    np.testing.assert_allclose(solver.residual_history(), o.hist, rtol=REL_TOL)
    # size-independent property: b = 1 and A symmetric persymmetric => x is symmetric about the middle
    x = solver.solution()
    assert rel_l2(x[::-1], x) <= 1e-12


# ---------------------------------------------------------------------------------- file mode
@pytest.mark.parametrize("n", [64, 200])
def test_file_mode_golden(solver, golden, golden_dir, n):
    """Files in the reference format -> load -> solve vs the solution file the reference CLI wrote."""
    g = [e for e in golden["file_mode"] if e["n"] == n][0]
    solver.load_matrix(os.path.join(golden_dir, f"spd_n{n}_A.bin"))
    solver.load_rhs(os.path.join(golden_dir, f"spd_n{n}_b.bin"))
    r = solver.solve(1000, 1e-9)
    x_ref = fileformat.read_vector(os.path.join(golden_dir, f"spd_n{n}_x.bin"))
    A = fileformat.read_matrix(os.path.join(golden_dir, f"spd_n{n}_A.bin"))
    b = fileformat.read_vector(os.path.join(golden_dir, f"spd_n{n}_b.bin"))
    # "same iteration count +-1": against the reference's own envelope over OMP_NUM_THREADS — the one recorded in the fixture by
    # tests/golden/make_golden.py AND the one measured now (the residual hovers around 1e-9 for several iterations on these
    # systems, so the unmodified reference itself moves by 2-4 iterations when only the thread count changes)
    spread = g["iters_by_omp_threads_1_to_8"]
    assert r.converged and min(spread) - 1 <= r.iterations <= max(spread) + 1, (r.iterations, spread)
    check_stopping_iteration(r, A, b, oracle.cg_solve(A, b, 1000, 1e-9), f"file_golden_n{n}")
    err = rel_l2(solver.solution(), x_ref)
    REPORT[f"file_golden_x_rel_l2_n{n}"] = err
    # the golden x stopped at g["iters"]; a run that stops on another iteration differs by one CG step near rel_err 1e-9 (~1e-10,
    # the unmodified reference differs from itself by 2.1e-10 this way); the sharp comparison is the matched one below
    assert err <= (X_TOL_STOPPED if r.iterations != g["iters"] else parity_util.x_tolerance(parity_util.reference_self_noise(A, b, g["iters"], oracle.cg_solve(A, b, g["iters"], 0.0).x)))
    check_matched_iterations(solver, A, b, g["iters"], f"file_golden_n{n}")


def test_file_mode_config5_n2048(solver, tmp_path):
    """BASELINE config 5: random_spd_system distribution, n = 2048, -i 1000 -e 1e-9 (about 351 iterations;
    the reference itself lands on 351..353 depending on OMP_NUM_THREADS)."""
    n = 2048
    A, b = random_spd.random_spd_system(n, 42)
    pa, pb, px = (str(tmp_path / f"{k}.bin") for k in "Abx")
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    solver.load_matrix(pa)
    solver.load_rhs(pb)
    r = solver.solve(1000, 1e-9)
    o = oracle.cg_solve(A, b, 1000, 1e-9, history=True)
    check_stopping_iteration(r, A, b, o, "file_n2048")
    x = solver.solution()
    err = rel_l2(x, o.x)
    REPORT["file_n2048_x_rel_l2"] = err
    assert err <= X_TOL_STOPPED  # may have stopped one iteration apart; the sharp x comparison is the matched one at the end
    # cond(A) ~ 1e3: summation-order differences grow along the recurrence (the reference differs from
    # itself the same way when OMP_NUM_THREADS changes), so the history is sharp early and loose late
    k = min(len(o.hist), r.iterations_run) - 5
    h = solver.residual_history()
    np.testing.assert_allclose(h[:40], o.hist[:40], rtol=1e-6)
    np.testing.assert_allclose(h[:k], o.hist[:k], rtol=0.5)
    # the true residual of the returned x agrees with what the solver reports
    true_rel = float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))
    assert true_rel < 2e-9
    # save -> bit-compatible file with a clean header (SURVEY 2.4 defect 1)
    solver.save_solution(px)
    assert fileformat.read_header(px) == (n, 1)
    assert np.array_equal(fileformat.read_vector(px), x)
    check_matched_iterations(solver, A, b, o.iters, "file_n2048")


def test_in_memory_system_host_and_device_pointers(lamcg, solver):
    """solve(A, b, x, ...) on caller-owned buffers: numpy (host) and torch (device) give the same bits."""
    import torch
    n = 777
    A, b = random_spd.random_spd_system(n, 5)
    solver.set_matrix(A)
    solver.set_rhs(b)
    r1 = solver.solve(1000, 1e-9)
    x1 = solver.solution()
    At = torch.from_numpy(A).cuda()
    bt = torch.from_numpy(b).cuda()
    solver.set_matrix(At)
    solver.set_rhs(bt)
    r2 = solver.solve(1000, 1e-9)
    x2 = solver.solution()
    assert r1.iterations == r2.iterations and np.array_equal(x1, x2)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    assert parity_util.iterations_within_one_of_reference(r1.iterations, A, b, 1000, 1e-9, o.iters)[0] and rel_l2(x1, o.x) <= X_TOL_STOPPED
    check_matched_iterations(solver, A, b, o.iters, "in_memory_n777")


# ------------------------------------------------------------------------- loop / determinism
def test_stream_and_graph_loops_are_bit_identical_and_reproducible(solver):
    n = 3000
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    results = []
    for mode in (1, 2, 2, 1):
        solver.set_option("loop_mode", mode)
        r = solver.solve(10000, 1e-9)
        results.append((r.iterations, r.rel_residual, solver.solution().copy()))
    for it, rel, x in results[1:]:
        assert it == results[0][0] and rel == results[0][1] and np.array_equal(x, results[0][2])


def test_graph_loop_times_every_gemv_like_the_stream_loop(solver):
    """loop_mode 2 + time_gemv 1: external event-record nodes around every K1 inside the captured chunks (two executables
    launched alternately so a chunk's events can be read while the next one runs).  Same bits as the untimed loops, a GEMV time
    that agrees with the stream loop's, and launches after convergence (no-ops) are not counted."""
    n = 20000
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    res = {}
    for name, opts in (("stream", {"loop_mode": 1, "time_gemv": 1}), ("graph_timed", {"loop_mode": 2, "time_gemv": 1}), ("graph", {"loop_mode": 2, "time_gemv": 0})):
        for k, v in opts.items():
            solver.set_option(k, v)
        solver.solve(100, 1e-9)
        r = solver.solve(100, 1e-9)
        res[name] = (r, solver.solution().copy())
    assert np.array_equal(res["stream"][1], res["graph_timed"][1]) and np.array_equal(res["graph"][1], res["graph_timed"][1])
    gs, gg = res["stream"][0].gemv_seconds / 100, res["graph_timed"][0].gemv_seconds / 100
    assert res["stream"][0].gemv_launches_timed == res["graph_timed"][0].gemv_launches_timed == 100
    assert res["graph"][0].gemv_seconds == 0.0 and gs > 0 and abs(gg - gs) <= 0.05 * gs, (gs, gg)
    assert gg <= res["graph_timed"][0].solve_seconds / 100                       # a part of the iteration, not more
    solver.set_option("loop_mode", 2)
    solver.set_option("time_gemv", 2)                                            # one GEMV per 16-iteration chunk
    r = solver.solve(100, 1e-9)
    assert r.gemv_launches_timed == 6 and abs(r.gemv_seconds / 6 - gs) <= 0.05 * gs   # chunks 0..5 have their sampled launch (8, 24, .. 88) below 100
    assert np.array_equal(solver.solution(), res["stream"][1])
    # converges at iteration 500 of n = 1000, in the middle of a 16-iteration chunk: only executed GEMVs are timed
    solver.set_option("loop_mode", 2)
    solver.set_option("time_gemv", 1)
    solver.generate_matrix(1000, 1000)
    solver.generate_rhs()
    r = solver.solve(10000, 1e-9)
    assert r.converged and r.iterations == 500 and 0 < r.gemv_seconds < r.solve_seconds


def test_graph_chunking_does_not_overshoot(solver):
    """max_iters not a multiple of the graph chunk, and convergence in the middle of a chunk."""
    n = 512
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    solver.set_option("loop_mode", 2)
    for chunk in (2, 16, 50):
        solver.set_option("chunk_iters", chunk)
        r = solver.solve(37, 1e-9)
        o = oracle.cg_solve_generated(n, 37, 1e-9)
        assert r.iterations == o.iters == 38 and r.iterations_run == 37
        assert math.isclose(r.rel_residual, o.rel, rel_tol=REL_TOL)
        r = solver.solve(10000, 1e-9)
        assert r.converged and r.iterations == 256 == r.iterations_run


def test_max_iters_zero_and_gemv_timing(solver):
    n = 300
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    r = solver.solve(0, 1e-9)
    assert not r.converged and r.iterations == 1 and r.iterations_run == 0  # loop variable after exit
    assert np.array_equal(solver.solution(), np.zeros(n))
    solver.set_option("time_gemv", 1)
    r = solver.solve(20, 1e-9)
    assert r.gemv_seconds > 0 and r.gemv_seconds <= r.solve_seconds
    assert solver.time_gemv(1, 3) > 0


# ------------------------------------------------------------------------------ error behaviour
def test_error_behaviour_matches_reference(lamcg, tmp_path, capsys):
    """bool returns + message on stderr, like OMP.hpp:98-118 / :151-155."""
    cg = lamcg.ConjugateGradient_B200(0, verbose=False)
    assert cg.load_matrix_from_file(str(tmp_path / "missing.bin")) is False
    assert "Cannot open" in capsys.readouterr().err
    rect = str(tmp_path / "rect.bin")
    fileformat.write_matrix(rect, np.ones((3, 4)))
    assert cg.load_matrix_from_file(rect) is False
    assert "Matrix has to be square" in capsys.readouterr().err
    A, b = random_spd.random_spd_system(32, 1)
    pa, pb, pbad, pwide = (str(tmp_path / f) for f in ("A.bin", "b.bin", "bad.bin", "wide.bin"))
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    fileformat.write_matrix(pbad, np.ones(31))
    fileformat.write_matrix(pwide, np.ones((32, 2)))
    assert cg.load_matrix_from_file(pa) is True
    assert cg.load_rhs_from_file(pbad) is False
    assert "Size of right hand side does not match the matrix" in capsys.readouterr().err
    assert cg.load_rhs_from_file(pwide) is False
    assert "does not contain a valid rhs" in capsys.readouterr().err
    assert cg.load_rhs_from_file(pb) is True
    assert cg.solve(1000, 1e-9) is True
    assert cg.solve(3, 1e-9) is False  # not converged -> false (OMP.hpp:86-90)
    assert cg.get_num_rows() == 32 and cg.get_num_cols() == 32
    assert cg.save_result_to_file(str(tmp_path / "nodir" / "x.bin")) is False
    cg.close()
    with pytest.raises(lamcg.LamcgError):
        s = lamcg.Solver(0)
        s.solve(10, 1e-9)  # no system yet


def test_truncated_matrix_file_is_rejected(solver, tmp_path, lamcg):
    p = str(tmp_path / "short.bin")
    with open(p, "wb") as f:
        np.array([100, 100], dtype=np.uint64).tofile(f)
        np.ones(50).tofile(f)
    with pytest.raises(lamcg.LamcgError) as e:
        solver.load_matrix(p)
    assert e.value.code == -3


# ------------------------------------------------------- persistent single-kernel loop (loop_mode 3)
@pytest.mark.parametrize("generation", [3, 4])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 147, 149, 1000, 1023, 1025, 2047, 2048, 2049, 3000, 4095, 4096, 5001, 10007])
def test_persistent_loop_generate_mode_vs_oracle(solver, n, generation):
    """The cooperative one-kernel loop (auto for n <= 16384, forced here): same exact iteration counts, residual history and x as
    the oracle; n around the CTA count exercises grids with 0/1/2 rows per CTA.  generation 4 = ONE exchange per iteration:
    all-gather of Ap as tagged words, p.Ap / r / r.r / beta / p computed redundantly by every CTA on register slices (n <= 4096,
    all three register widths: lda <= 1024 / 2048 / 4096; odd n exercises the single-entry tail of the gather; auto below n = 4081);
    generation 3 = K1's streaming row sweep inside the loop with two scalar exchanges (auto from there up)."""
    if generation == 4 and n > 4096:
        pytest.skip("the fourth-generation kernel holds p in registers: n <= 4096")
    max_iters = 10000 if n <= 4096 else 300
    solver.set_option("loop_mode", 3)
    solver.set_option("persist_variant", generation)
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    r = solver.solve(max_iters, 1e-9)
    assert r.kernel_launches == 1
    o = oracle.cg_solve_generated(n, max_iters, 1e-9, history=True)
    assert r.iterations == o.iters and bool(r.converged) == o.converged and r.iterations_run == min(o.iters, max_iters)
    err = rel_l2(solver.solution(), o.x)
    REPORT[f"persistent_gen_x_rel_l2_n{n}"] = err
    # runs that end by finite termination at ceil(n/2) finish with a pure-rounding step (see LONG_RUN note)
    # ... and for odd n that last step is ill conditioned (cond ~ 0.4 n^2): at n = 4095 the UNMODIFIED reference differs from
    # itself by 1.79e-10 between OMP_NUM_THREADS = 1 and 8 (2048 iterations both; measured with oracle.ref_gen_solve), we
    # differ from it by 1.82e-10 -> the north_star tolerance is widened with n^2 there (5e-10 at n = 4095)
    assert err <= (X_TOL_GEN if (o.iters <= LONG_RUN and not o.converged) else max(X_TOL, 3e-17 * n * n))
    h = solver.residual_history()
    big = o.hist > 1e-9
    np.testing.assert_allclose(h[big], o.hist[big], rtol=REL_TOL)
    # a second solve on the same handle restarts from x0 = 0 and reproduces the same bits
    x1 = solver.solution().copy()
    r2 = solver.solve(max_iters, 1e-9)
    assert r2.iterations == r.iterations and np.array_equal(solver.solution(), x1)


def test_persistent_loop_agrees_with_graph_loop_on_spd(solver):
    n = 1536
    A, b = random_spd.random_spd_system(n, 21)
    solver.set_matrix(A)
    solver.set_rhs(b)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    env = parity_util.reference_iteration_envelope(A, b, 1000, 1e-9, o.iters)  # the unmodified reference over OMP_NUM_THREADS, measured now
    out = {}
    for mode, gen in ((2, 0), (3, 3), (3, 4)):
        solver.set_option("loop_mode", mode)
        solver.set_option("persist_variant", gen)
        r = solver.solve(1000, 1e-9)
        assert r.converged and env[0] - 1 <= r.iterations <= env[1] + 1, (r.iterations, env)
        out[mode, gen] = (r.iterations, solver.solution().copy(), r.iterations_run / r.solve_seconds)
        assert rel_l2(out[mode, gen][1], o.x) <= X_TOL_STOPPED
    REPORT["spd_n1536_it_per_s_graph_vs_persistent_gen3_gen4"] = [out[2, 0][2], out[3, 3][2], out[3, 4][2]]
    for gen in (3, 4):
        solver.set_option("loop_mode", 3)
        solver.set_option("persist_variant", gen)
        check_matched_iterations(solver, A, b, o.iters, f"persistent_gen{gen}_spd_n1536")
    r = solver.solve(0, 1e-9)
    assert not r.converged and r.iterations == 1 and r.iterations_run == 0


@pytest.mark.parametrize("loop_mode", [1, 2, 3])
def test_non_finite_systems_report_like_the_reference(solver, loop_mode):
    """b = 0 gives rr/bb = 0/0: the reference iterates on NaNs to the end and reports max_iters+1 and nan
    (the `10001,-nan` rows of TESTS/BEST_RESULTS:114).  Same report, but the loop stops at once."""
    n = 64
    solver.set_option("loop_mode", loop_mode)
    solver.set_matrix(oracle.generate_matrix(n))
    solver.set_rhs(np.zeros(n))
    r = solver.solve(500, 1e-9)
    o = oracle.cg_solve(oracle.generate_matrix(n), np.zeros(n), 500, 1e-9)
    assert not r.converged and not o.converged
    assert r.iterations == o.iters == 501 and math.isnan(r.rel_residual) and math.isnan(o.rel)
    assert r.numerical_breakdown == 1 and r.iterations_run == 1
    # a healthy system afterwards on the same handle
    solver.set_rhs(np.ones(n))
    r = solver.solve(500, 1e-9)
    assert r.converged and r.numerical_breakdown == 0 and r.iterations == 32


def test_refused_cooperative_launch_falls_back_to_the_graph_loop(solver, lamcg):
    """The one-kernel loop needs all its CTAs resident at once.  When the device refuses the cooperative launch (simulated by the
    debug_persist_fail option, a test hook) a solve whose loop was chosen by size runs through the graph loop with the same
    result and leaves no error message behind; a solve that asked for loop_mode 3 explicitly reports the failure."""
    n = 1000
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    o = oracle.cg_solve_generated(n, 10000, 1e-9)
    r = solver.solve(10000, 1e-9)
    assert r.kernel_launches == 1 and r.iterations == o.iters            # auto: the one-kernel loop
    solver.set_option("debug_persist_fail", 1)
    r = solver.solve(10000, 1e-9)
    assert r.kernel_launches > 1 and r.converged and r.iterations == o.iters
    assert rel_l2(solver.solution(), o.x) <= X_TOL
    assert (lamcg.lib().lamcg_last_error(solver._h) or b"") == b""      # the fallback succeeded: no stale error text
    solver.set_option("loop_mode", 3)
    with pytest.raises(lamcg.LamcgError) as e:
        solver.solve(10000, 1e-9)
    assert e.value.code == -2 and "cooperative launch" in e.value.message
    solver.set_option("debug_persist_fail", 0)
    r = solver.solve(10000, 1e-9)
    assert r.kernel_launches == 1 and r.iterations == o.iters


@pytest.mark.parametrize("n,persist_variant,too_few,enough", [(9000, 0, 16, 18), (9000, 3, 16, 18), (3000, 4, 5, 6), (3000, 0, 5, 8)])
def test_one_kernel_loop_on_a_device_with_few_sms(solver, lamcg, n, persist_variant, too_few, enough):
    """Both one-kernel loops own one row per thread (512 per CTA).  On a device that offers few SMs (MIG slice, MPS limit;
    simulated with persist_grid) a system may need more: the size-chosen loop must fall back to the graph loop and still be
    right, an explicit loop_mode 3 must be refused — never a silent wrong x (round-1 ADVICE); with just enough CTAs the kernel
    must be right on that small grid (hundreds of rows per CTA, most of them streamed from L2)."""
    solver.generate_matrix(n, n)
    solver.generate_rhs()
    o = oracle.cg_solve_generated(n, 60, 1e-9)
    solver.set_option("persist_variant", persist_variant)
    solver.set_option("persist_grid", too_few)                           # e.g. ceil(9000 / 16) = 563 rows per CTA > 512
    r = solver.solve(60, 1e-9)
    assert r.kernel_launches > 1 and r.iterations == o.iters
    assert rel_l2(solver.solution(), o.x) <= X_TOL_GEN
    solver.set_option("loop_mode", 3)
    with pytest.raises(lamcg.LamcgError) as e:
        solver.solve(60, 1e-9)
    assert e.value.code == -1 and "rows per CTA" in e.value.message
    solver.set_option("persist_grid", enough)                            # e.g. 500 rows per CTA: fits, and must be right on 18 CTAs
    r = solver.solve(60, 1e-9)
    assert r.kernel_launches == 1 and r.iterations == o.iters
    assert rel_l2(solver.solution(), o.x) <= X_TOL_GEN
