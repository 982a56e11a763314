"""The reference-compatible C++ drivers (LAM.hpp -> ConjugateGradient_B200 -> C ABI) on the GPU:
CSV contract of the getopt driver, positional driver, exit codes, file-format compatibility with the
reference's own CLI (oracle/_ref/test_CG_CPU_OMP.out reads what we write and vice versa)."""
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
TEST_DIR = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "test")
GETOPT = os.path.join(TEST_DIR, "test_CG_MultiGPUS_CUDA_MPI.out")        # 9-field CSV line (test_CG_CPU_MPI_OMP.cpp:201-203)
GETOPT_NCCL = os.path.join(TEST_DIR, "test_CG_MultiGPUS_CUDA_NCCL.out") # 10 fields: + communicator init after io_s (NCCL.cu:329-334)
POSITIONAL = os.path.join(TEST_DIR, "test_CG_single_GPU.out")


def run(cmd, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=timeout)


def test_generate_mode_csv_line_matches_reference_rows(golden, tmp_path):
    """Same 9 CSV fields as test_CG_CPU_MPI_OMP.out; iteration count and residual equal the reference's."""
    for g in golden["generate_mode_cli"]:
        res = run([GETOPT, "-s", str(g["n"]), "-i", str(g["max_iters"]), "-e", "1e-9", "-o", str(tmp_path / "sol.bin")])
        assert res.returncode == 0, res.stderr
        f = res.stdout.strip().split(",")
        assert len(f) == g["csv_fields"] == 9, res.stdout
        assert int(f[0]) == g["n"] and int(f[1]) == 1 and int(f[2]) == 1
        assert int(f[6]) == g["iters"]
        assert math.isclose(float(f[7]), g["rel_printed"], rel_tol=2e-5)
        float(f[3]); float(f[4]); float(f[5])
        assert f[8] == str(int(float(f[8])))  # whole seconds in generate mode, like the reference
        x = fileformat.read_vector(str(tmp_path / "sol.bin"))
        o = oracle.cg_solve_generated(g["n"], g["max_iters"], 1e-9)
        assert np.linalg.norm(x - o.x) / np.linalg.norm(o.x) <= 1e-12  # we save x (the reference saves b: defect 2)


def test_generate_mode_with_the_matrix_held_in_fp32_by_environment(golden, tmp_path):
    """LAMCG_MATRIX_F32=1 switches the unmodified driver to the mixed storage (option matrix_f32); generate mode's 0 / 1 / 2 are fp32
    numbers, so iteration count, printed residual and x are those of the reference all the same."""
    g = [e for e in golden["generate_mode_cli"] if e["n"] == 10000 and e["max_iters"] == 15][0]
    res = run([GETOPT, "-s", str(g["n"]), "-i", str(g["max_iters"]), "-e", "1e-9", "-o", str(tmp_path / "sol.bin")], env={"LAMCG_MATRIX_F32": "1"})
    assert res.returncode == 0, res.stderr
    f = res.stdout.strip().split(",")
    assert len(f) == 9 and int(f[6]) == g["iters"] and math.isclose(float(f[7]), g["rel_printed"], rel_tol=2e-5)
    x = fileformat.read_vector(str(tmp_path / "sol.bin"))
    o = oracle.cg_solve_generated(g["n"], g["max_iters"], 1e-9)
    assert np.linalg.norm(x - o.x) / np.linalg.norm(o.x) <= 1e-12


@pytest.mark.parametrize("exe,fields", [(GETOPT, 9), (GETOPT_NCCL, 10)])
def test_csv_field_count_per_executable(exe, fields, tmp_path):
    """The reference ships two getopt GPU executables whose CSV lines differ by one field: test_CG_MultiGPUS_CUDA_MPI.out prints
    n,ranks,threads,io_s,avg_gemv_s,avg_iter_s,iters,rel_err,total_s and test_CG_MultiGPUS_CUDA_NCCL.out inserts the communicator
    init seconds after io_s (ConjugateGradient_MultiGPUS_CUDA_NCCL.cu:329-334, e.g. TESTS/BEST_RESULTS:420
    `10000,1,1,0.986,1.54201,0.000698578,0.000810044,323,9.424e-10,1.832`).  Same here, in generate and in file mode."""
    n, k = 2048, 15
    o = oracle.cg_solve_generated(n, k, 1e-9)
    res = run([exe, "-s", str(n), "-i", str(k), "-e", "1e-9", "-o", str(tmp_path / "sol.bin")])
    assert res.returncode == 0, res.stderr
    assert res.stdout.endswith("\n") and res.stdout.count("\n") == 1      # exactly one line, closed by std::endl
    f = res.stdout.strip().split(",")
    assert len(f) == fields, res.stdout
    shift = fields - 9
    assert [int(f[0]), int(f[1]), int(f[2])] == [n, 1, 1]
    assert all(float(v) >= 0.0 for v in f[3:6 + shift])                   # io_s [, comm_init_s], avg_gemv_s, avg_iter_s
    assert int(f[6 + shift]) == o.iters == k + 1
    assert math.isclose(float(f[7 + shift]), o.rel, rel_tol=2e-5)
    assert f[8 + shift] == str(int(float(f[8 + shift])))                   # whole seconds in generate mode
    A, b = random_spd.random_spd_system(64, 5)
    pa, pb = str(tmp_path / "A.bin"), str(tmp_path / "b.bin")
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    res = run([exe, "-A", pa, "-b", pb, "-o", str(tmp_path / "x.bin"), "-i", "1000", "-e", "1e-9"])
    assert res.returncode == 0, res.stderr
    f = res.stdout.strip().split(",")
    assert len(f) == fields and int(f[0]) == 64, res.stdout
    # verbose mode suppresses the driver-side fields but the class-side ones (n, and avg_gemv..rel) still appear (MPI_OMP.hpp:203-205,122-127)
    res = run([exe, "-s", "64", "-v", "-o", str(tmp_path / "v.bin")])
    assert res.returncode == 0 and "Number of threads: 1" in res.stdout


def test_file_mode_both_drivers_and_reference_cli_interop(tmp_path):
    n = 300
    A, b = random_spd.random_spd_system(n, 9)
    pa, pb, px, px2, pxr = (str(tmp_path / f) for f in ("A.bin", "b.bin", "x.bin", "x2.bin", "xr.bin"))
    fileformat.write_matrix(pa, A)
    fileformat.write_matrix(pb, b)
    o = oracle.cg_solve(A, b, 1000, 1e-9)
    # positional driver (test_CG_single_GPU / test_CG_CPU_OMP convention)
    res = run([POSITIONAL, pa, pb, px, "1000", "1e-9"])
    assert res.returncode == 0, res.stderr
    assert "Converged in" in res.stdout and "Finished successfully" in res.stdout
    x = fileformat.read_vector(px)
    assert fileformat.read_header(px) == (n, 1)
    # runs that stop on different iterations differ by ~1e-10 (the unmodified reference moves 258..262 on this
    # system with OMP_NUM_THREADS and differs from itself by 2.1e-10) -> loose here, sharp at matched count below
    assert np.linalg.norm(x - o.x) / np.linalg.norm(o.x) <= 1e-9
    res = run([POSITIONAL, pa, pb, px, str(o.iters), "0"])  # rel_error 0: exactly o.iters iterations on both sides
    assert res.returncode == 0 and "Did not converge in %d iterations" % o.iters in res.stdout
    om = oracle.cg_solve(A, b, o.iters, 0.0)
    xm = fileformat.read_vector(px)
    assert parity_util.rel_l2(xm, om.x) <= parity_util.x_tolerance(parity_util.reference_self_noise(A, b, o.iters, om.x))
    res = run([POSITIONAL, pa, pb, px, "1000", "1e-9"])
    x = fileformat.read_vector(px)
    # getopt driver, file mode: fractional seconds in the last field
    res = run([GETOPT, "-A", pa, "-b", pb, "-o", px2, "-i", "1000", "-e", "1e-9"])
    assert res.returncode == 0, res.stderr
    f = res.stdout.strip().split(",")
    assert len(f) == 9 and int(f[0]) == n and parity_util.iterations_within_one_of_reference(int(f[6]), A, b, 1000, 1e-9, o.iters)[0]
    assert np.array_equal(fileformat.read_vector(px2), x)  # same library, same bits
    # the reference CLI accepts our files as its input (and we solve what it solves)
    if os.path.exists(oracle.REF_TEST_OMP):
        fileformat.write_matrix(str(tmp_path / "x_as_rhs.bin"), x)  # a vector file we wrote, used as an rhs by the reference
        r = run([oracle.REF_TEST_OMP, pa, str(tmp_path / "x_as_rhs.bin"), pxr, "5", "1e-9"], env={"OMP_NUM_THREADS": "1"})
        assert r.returncode == 0, r.stderr


def test_cli_errors_and_exit_codes(tmp_path):
    pa, pb = str(tmp_path / "A.bin"), str(tmp_path / "b.bin")
    res = run([POSITIONAL, str(tmp_path / "nope.bin"), pb, str(tmp_path / "x.bin")])
    assert res.returncode == 1 and "Failed to read matrix" in res.stderr
    fileformat.write_matrix(pa, np.eye(4))
    fileformat.write_matrix(pb, np.ones(5))
    res = run([POSITIONAL, pa, pb, str(tmp_path / "x.bin")])
    assert res.returncode == 2 and "Failed to read right hand side" in res.stderr
    fileformat.write_matrix(pb, np.ones(4))
    res = run([POSITIONAL, pa, pb, str(tmp_path / "no_such_dir" / "x.bin")])
    assert res.returncode == 6 and "Failed to save solution" in res.stderr
    res = run([GETOPT, "-s", "100", "-A", pa])
    assert res.returncode == 1 and "cannot be used with" in res.stderr
    res = run([GETOPT, "-A", pa, "-s", "100"])
    assert res.returncode == 1 and "cannot be used with -s" in res.stderr
    res = run([GETOPT, "-h"])
    assert res.returncode == 0 and "-s <int>" in res.stdout
    res = run([GETOPT, "-s", "64", "-v", "-o", str(tmp_path / "v.bin")])
    assert res.returncode == 0 and "Finished successfully" in res.stdout
