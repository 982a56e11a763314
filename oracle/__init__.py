"""CPU oracle for the dense-CG hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker (or, for the bench,
as the timed CPU baseline).  The product path (the ``lamcg`` CUDA library and its host classes)
never imports, links or executes anything under ``oracle/``.

Parity status: PINNED — see ``cg_oracle.c`` header and ``tests/test_oracle.py``.

Contents
--------
* ``cg_oracle.c``      our C restatement of the reference loop (-> ``liboracle_cg.so``)
* ``ref_harness.cpp``  C-ABI window on the unmodified reference classes (-> ``_ref/libref_harness.so``)
* ``ref_gpu_harness.cu`` command-line window on the unmodified reference GPU classes (-> ``_ref/ref_gpu_{single,multi}.out``);
  ``_ref/test_CG_single_GPU.out`` / ``test_CG_MultiGPUS_CUDA.out`` are the reference's own GPU drivers, unmodified
* ``mpi_shim/mpi.h``   1-rank MPI stand-in so the reference compiles without MPI
* ``fileformat.py``    numpy restatement of the binary matrix/rhs/solution format
* ``random_spd.py``    numpy restatement of ``random_spd_system.cpp``'s SPD distribution
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle_cg.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HARNESS_SO = os.path.join(REF_DIR, "libref_harness.so")
REF_TEST_OMP = os.path.join(REF_DIR, "test_CG_CPU_OMP.out")
REF_TEST_MPI_OMP = os.path.join(REF_DIR, "test_CG_CPU_MPI_OMP.out")
REF_TEST_SINGLE_GPU = os.path.join(REF_DIR, "test_CG_single_GPU.out")
REF_TEST_MULTI_GPU = os.path.join(REF_DIR, "test_CG_MultiGPUS_CUDA.out")
REF_GPU_HARNESS = {"single": os.path.join(REF_DIR, "ref_gpu_single.out"), "multi": os.path.join(REF_DIR, "ref_gpu_multi.out")}

_c_double_p = ctypes.POINTER(ctypes.c_double)


def build(ref: bool | None = None) -> None:
    """Compile the checkers (``make -C oracle``).  ``ref=None`` builds ``_ref`` only when the
    reference sources are present (i.e. in the build container, never on the GPU box)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir("/root/reference/challenge/main")
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref", "refgpu"], check=True)


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_c_double_p)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = ctypes.CDLL(ORACLE_SO)
        L.oracle_dot.restype = ctypes.c_double
        L.oracle_dot.argtypes = [_c_double_p, _c_double_p, ctypes.c_size_t]
        L.oracle_axpby.restype = None
        L.oracle_axpby.argtypes = [ctypes.c_double, _c_double_p, ctypes.c_double, _c_double_p, ctypes.c_size_t]
        L.oracle_gemv.restype = None
        L.oracle_gemv.argtypes = [ctypes.c_double, _c_double_p, _c_double_p, ctypes.c_double, _c_double_p,
                                  ctypes.c_size_t, ctypes.c_size_t]
        L.oracle_partition.restype = None
        L.oracle_partition.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_generate_matrix.restype = None
        L.oracle_generate_matrix.argtypes = [_c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
        L.oracle_generate_rhs.restype = None
        L.oracle_generate_rhs.argtypes = [_c_double_p, ctypes.c_size_t]
        L.oracle_cg_solve.restype = ctypes.c_int
        L.oracle_cg_solve.argtypes = [_c_double_p, _c_double_p, _c_double_p, ctypes.c_size_t, ctypes.c_int,
                                      ctypes.c_double, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_double), _c_double_p, ctypes.c_size_t]
        L.oracle_cg_solve_generated.restype = ctypes.c_int
        L.oracle_cg_solve_generated.argtypes = [ctypes.c_size_t, _c_double_p, ctypes.c_int, ctypes.c_double,
                                                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double),
                                                _c_double_p, ctypes.c_size_t]
        L.oracle_gemv_generated.restype = None
        L.oracle_gemv_generated.argtypes = [_c_double_p, _c_double_p, ctypes.c_size_t]
        L.oracle_rand_fill.restype = None
        L.oracle_rand_fill.argtypes = [_c_double_p, ctypes.c_size_t, ctypes.c_int]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------ primitives
def dot(x: np.ndarray, y: np.ndarray) -> float:
    return float(lib().oracle_dot(_dp(x), _dp(y), x.size))


def axpby(alpha: float, x: np.ndarray, beta: float, y: np.ndarray) -> None:
    lib().oracle_axpby(alpha, _dp(x), beta, _dp(y), x.size)


def gemv(A: np.ndarray, x: np.ndarray, alpha: float = 1.0, beta: float = 0.0, y: np.ndarray | None = None) -> np.ndarray:
    rows, cols = A.shape
    if y is None:
        y = np.zeros(rows)
    lib().oracle_gemv(alpha, _dp(A), _dp(x), beta, _dp(y), rows, cols)
    return y


def gemv_generated(p: np.ndarray) -> np.ndarray:
    out = np.zeros_like(p)
    lib().oracle_gemv_generated(_dp(p), _dp(out), p.size)
    return out


def partition(n: int, nranks: int, rank: int) -> tuple[int, int]:
    """(local_rows, row_offset) — ref MPI_OMP.hpp:175-184."""
    rows, off = ctypes.c_size_t(), ctypes.c_size_t()
    lib().oracle_partition(n, nranks, rank, ctypes.byref(rows), ctypes.byref(off))
    return int(rows.value), int(off.value)


def generate_matrix(n: int, local_rows: int | None = None, offset: int = 0) -> np.ndarray:
    local_rows = n if local_rows is None else local_rows
    A = np.empty((local_rows, n))
    lib().oracle_generate_matrix(_dp(A), local_rows, n, offset)
    return A


def generate_rhs(n: int) -> np.ndarray:
    b = np.empty(n)
    lib().oracle_generate_rhs(_dp(b), n)
    return b


class Result:
    """What the reference reports for one solve: iteration count (max_iters+1 when not converged,
    MPI_OMP.hpp:125), relative residual, converged flag, x and the per-iteration residual history."""

    def __init__(self, converged, iters, rel, x, hist=None, seconds=None):
        self.converged, self.iters, self.rel, self.x, self.hist, self.seconds = converged, iters, rel, x, hist, seconds

    def __repr__(self):
        return f"Result(converged={self.converged}, iters={self.iters}, rel={self.rel:.6e})"


def cg_solve(A: np.ndarray, b: np.ndarray, max_iters: int, rel_error: float, history: bool = False) -> Result:
    n = b.size
    x = np.zeros(n)
    it, rel = ctypes.c_int(), ctypes.c_double()
    hist = np.zeros(max_iters) if history else None
    rc = lib().oracle_cg_solve(_dp(A), _dp(b), _dp(x), n, max_iters, rel_error, ctypes.byref(it), ctypes.byref(rel),
                               _dp(hist) if history else None, max_iters if history else 0)
    assert rc >= 0
    if history:
        hist = hist[: min(it.value, max_iters)]
    return Result(bool(rc), it.value, rel.value, x, hist)


def cg_solve_generated(n: int, max_iters: int, rel_error: float, history: bool = False) -> Result:
    """Generate-mode solve (A = tridiag(1,2,1) stored dense in the reference, b = 1) in O(n) memory,
    bit-identical to the dense oracle (see cg_oracle.c: matvec_generated)."""
    x = np.zeros(n)
    it, rel = ctypes.c_int(), ctypes.c_double()
    hist = np.zeros(max_iters) if history else None
    rc = lib().oracle_cg_solve_generated(n, _dp(x), max_iters, rel_error, ctypes.byref(it), ctypes.byref(rel),
                                         _dp(hist) if history else None, max_iters if history else 0)
    assert rc >= 0
    if history:
        hist = hist[: min(it.value, max_iters)]
    return Result(bool(rc), it.value, rel.value, x, hist)


def cg_solve_f32(A, b, max_iters: int, rel_error: float) -> Result:
    """fp32 restatement (every variable a float), A = None selects the generate-mode matrix."""
    L = lib()
    fp = ctypes.POINTER(ctypes.c_float)
    L.oracle_cg_solve_f32.restype = ctypes.c_int
    L.oracle_cg_solve_f32.argtypes = [fp, fp, fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_float)]
    b = np.ascontiguousarray(b, dtype=np.float32)
    n = b.size
    x = np.zeros(n, dtype=np.float32)
    it, rel = ctypes.c_int(), ctypes.c_float()
    Ap = None
    if A is not None:
        A = np.ascontiguousarray(A, dtype=np.float32)
        Ap = A.ctypes.data_as(fp)
    rc = L.oracle_cg_solve_f32(Ap, b.ctypes.data_as(fp), x.ctypes.data_as(fp), n, max_iters, rel_error, ctypes.byref(it), ctypes.byref(rel))
    assert rc >= 0
    return Result(bool(rc), it.value, float(rel.value), x)


def ref_omp_solve_f32(A, b, max_iters: int, rel_error: float, threads: int | None = None) -> Result:
    """The unmodified reference's ConjugateGradient_CPU_OMP<float>::solve on an in-memory system."""
    R = ref()
    if threads:
        R.ref_set_threads(threads)
    fp = ctypes.POINTER(ctypes.c_float)
    R.ref_omp_solve_f32.restype = ctypes.c_int
    R.ref_omp_solve_f32.argtypes = [fp, fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, fp, ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_float)]
    A = np.ascontiguousarray(A, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    n = b.size
    x = np.zeros(n, dtype=np.float32)
    it, rel = ctypes.c_int(), ctypes.c_float()
    rc = R.ref_omp_solve_f32(A.ctypes.data_as(fp), b.ctypes.data_as(fp), n, max_iters, rel_error, x.ctypes.data_as(fp),
                             ctypes.byref(it), ctypes.byref(rel))
    assert rc >= 0, "could not parse the reference's output"
    return Result(bool(rc), it.value, float(rel.value), x)


def rand_fill(count: int, seed: int) -> np.ndarray:
    out = np.empty(count)
    lib().oracle_rand_fill(_dp(out), count, seed)
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ------------------------------------------------------------------ the real reference (_ref)
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_HARNESS_SO)


def ref() -> ctypes.CDLL:
    """oracle/_ref/libref_harness.so: the unmodified reference classes (built by `make ref`)."""
    global _ref
    if _ref is None:
        R = ctypes.CDLL(REF_HARNESS_SO)
        R.ref_num_threads.restype = ctypes.c_int
        R.ref_set_threads.argtypes = [ctypes.c_int]
        R.ref_gen_solve.restype = ctypes.c_int
        R.ref_gen_solve.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_double, _c_double_p,
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        R.ref_omp_solve.restype = ctypes.c_int
        R.ref_omp_solve.argtypes = [_c_double_p, _c_double_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double,
                                    _c_double_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        _ref = R
    return _ref


def ref_gen_solve(n: int, max_iters: int, rel_error: float, threads: int | None = None) -> Result:
    R = ref()
    if threads:
        R.ref_set_threads(threads)
    x = np.zeros(n)
    it, rel, secs, gsecs = ctypes.c_int(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    rc = R.ref_gen_solve(n, max_iters, rel_error, _dp(x), ctypes.byref(it), ctypes.byref(rel),
                         ctypes.byref(secs), ctypes.byref(gsecs))
    assert rc >= 0, "could not parse the reference's output"
    res = Result(bool(rc), it.value, rel.value, x, None, secs.value)
    res.gen_seconds = gsecs.value
    return res


class RefGenSystem:
    """The unmodified reference's generate-mode system kept alive between solves (``ref_gen_open`` in ref_harness.cpp):
    generate once, ``solve(k)`` as often as needed.  Raises MemoryError when the reference cannot allocate the n x n matrix."""

    def __init__(self, n: int, threads: int | None = None):
        R = ref()
        if threads:
            R.ref_set_threads(threads)
        R.ref_gen_open.restype = ctypes.c_void_p
        R.ref_gen_open.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_double)]
        R.ref_gen_solve_again.restype = ctypes.c_int
        R.ref_gen_solve_again.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, _c_double_p, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        R.ref_gen_close.restype = None
        R.ref_gen_close.argtypes = [ctypes.c_void_p]
        g = ctypes.c_double()
        self.n, self._R = n, R
        self._h = R.ref_gen_open(n, ctypes.byref(g))
        self.gen_seconds = g.value
        if not self._h:
            raise MemoryError(f"the reference could not generate its {n} x {n} system")

    def solve(self, max_iters: int, rel_error: float, want_x: bool = True) -> Result:
        x = np.zeros(self.n) if want_x else None
        it, rel, secs = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
        rc = self._R.ref_gen_solve_again(self._h, max_iters, rel_error, _dp(x) if want_x else None, ctypes.byref(it), ctypes.byref(rel),
                                         ctypes.byref(secs))
        assert rc >= 0, "could not parse the reference's output"
        return Result(bool(rc), it.value, rel.value, x, None, secs.value)

    def close(self) -> None:
        if self._h:
            self._R.ref_gen_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


def ref_omp_solve(A: np.ndarray, b: np.ndarray, max_iters: int, rel_error: float, threads: int | None = None) -> Result:
    R = ref()
    if threads:
        R.ref_set_threads(threads)
    n = b.size
    x = np.zeros(n)
    it, rel, secs = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
    rc = R.ref_omp_solve(_dp(A), _dp(b), n, max_iters, rel_error, _dp(x), ctypes.byref(it), ctypes.byref(rel),
                         ctypes.byref(secs))
    assert rc >= 0, "could not parse the reference's output"
    return Result(bool(rc), it.value, rel.value, x, None, secs.value)


# ------------------------------------------------------------------ the reference's GPU classes on this GPU (_ref, refgpu)
def ref_gpu_available(variant: str = "single") -> bool:
    return os.path.exists(REF_GPU_HARNESS[variant])


def ref_gpu_solve(variant: str, max_iters, rel_error: float, n: int | None = None, A_path: str | None = None,
                  b_path: str | None = None, x_path: str | None = None, timeout: float = 600.0) -> list[dict]:
    """Run the unmodified reference GPU class (``single`` = ConjugateGradient_GPU_CUDA, ``multi`` =
    ConjugateGradient_MultiGPUS_CUDA) in its own process through ``ref_gpu_harness.cu``: generate mode when ``n`` is
    given, file mode otherwise.  ``max_iters`` may be a list: one solve per entry on the same resident host system
    (the reference's solve() uploads A every call, so loop time = difference between two entries).  Returns one dict
    per solve (iters as the class prints them, rel, wall seconds of solve()); x of the last solve goes to ``x_path``."""
    import json
    ks = [max_iters] if isinstance(max_iters, int) else list(max_iters)
    exe = REF_GPU_HARNESS[variant]
    if n is not None:
        cmd = [exe, "gen", str(n)]
    else:
        cmd = [exe, "file", A_path, b_path]
    cmd += [repr(float(rel_error)), x_path or "-"] + [str(k) for k in ks]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout).stdout
    return [json.loads(line) for line in out.splitlines() if line.startswith("{")]
