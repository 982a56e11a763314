"""B200-native dense Conjugate Gradient — Python host mirror of the reference's solver interface.

The product is ``liblamcg.so`` (hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/lamcg.h``).  This module is only a thin ctypes binding plus ``ConjugateGradient_B200``,
a class with the same method names, argument meaning and bool/print behaviour as the reference's
``LAM::ConjugateGradient<T>`` hierarchy (challenge/main/LAM/src/ConjugateGradient.hpp:9-28,
CPU/ConjugateGradient_CPU_MPI_OMP.hpp:19-69), so tests read like runs of the reference drivers.
The reference itself is C++; the C++ twin of this class is ``LAM/src/B200/ConjugateGradient_B200.hpp``.

There is no CPU fallback: importing works anywhere (so the CPU test-suite can check the ABI), but
creating a solver without a CUDA device raises ``LamcgError``.

The directory name is not a valid Python identifier; import it with
``importlib.import_module("2024-eumaster4hpc-student-challenge_b200")`` or via ``lamcg_b200.py``
at the repository root.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

from . import launch  # noqa: F401  (host-side multi-rank plumbing)

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.environ.get("LAMCG_LIB") or os.path.join(HERE, "liblamcg.so")  # LAMCG_LIB: an alternative build of the same library (A/B measurements)
HEADER_PATH = os.path.join(REPO, "include", "lamcg.h")

NCCL_ID_BYTES = 128
PEER_HANDLE_BYTES = 128

STATUS = {0: "OK", -1: "INVALID", -2: "CUDA", -3: "IO", -4: "SHAPE", -5: "NOMEM", -6: "COMM", -7: "STATE", -8: "DEVICE"}


class LamcgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lamcg error {code} ({STATUS.get(code, '?')}): {message}")
        self.code = code
        self.message = message


class lamcg_result(ctypes.Structure):
    _fields_ = [("converged", ctypes.c_int), ("iterations", ctypes.c_int), ("rel_residual", ctypes.c_double),
                ("solve_seconds", ctypes.c_double), ("gemv_seconds", ctypes.c_double),
                ("iterations_run", ctypes.c_int), ("kernel_launches", ctypes.c_int), ("numerical_breakdown", ctypes.c_int),
                ("gemv_launches_timed", ctypes.c_int)]


class lamcg_info(ctypes.Structure):
    _fields_ = [("n", ctypes.c_size_t), ("local_rows", ctypes.c_size_t), ("row_offset", ctypes.c_size_t),
                ("lda", ctypes.c_size_t), ("rank", ctypes.c_int), ("nranks", ctypes.c_int), ("device", ctypes.c_int),
                ("sm_count", ctypes.c_int), ("comm_mode", ctypes.c_int), ("has_matrix", ctypes.c_int),
                ("has_rhs", ctypes.c_int), ("gemv_variant", ctypes.c_int), ("gemv_grid", ctypes.c_int),
                ("gemv_block", ctypes.c_int), ("gemv_smem_bytes", ctypes.c_int), ("dtype", ctypes.c_int),
                ("ingest_threads", ctypes.c_int), ("ingest_chunks", ctypes.c_int), ("matrix_elem_bytes", ctypes.c_int),
                ("matrix_f32_inexact", ctypes.c_ulonglong), ("matrix_f32_overflow", ctypes.c_ulonglong)]


def build(verbose: bool = False) -> str:
    """Compile liblamcg.so in tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", HERE, "lib"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building liblamcg.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None
_vp, _cp, _dp = ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes): must list every symbol include/lamcg.h declares (tests check that).
_SIGNATURES = {
    "lamcg_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    "lamcg_create_ranked": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lamcg_create_typed": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lamcg_destroy": (None, [_vp]),
    "lamcg_last_error": (_cp, [_vp]),
    "lamcg_version": (_cp, []),
    "lamcg_set_option": (ctypes.c_int, [_vp, _cp, ctypes.c_longlong]),
    "lamcg_get_info": (ctypes.c_int, [_vp, ctypes.POINTER(lamcg_info)]),
    "lamcg_comm_nccl_unique_id": (ctypes.c_int, [_vp]),
    "lamcg_comm_init_nccl": (ctypes.c_int, [_vp, _vp]),
    "lamcg_comm_peer_export": (ctypes.c_int, [_vp, ctypes.c_size_t, _vp]),
    "lamcg_comm_init_peer": (ctypes.c_int, [_vp, _vp]),
    "lamcg_generate_matrix": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_size_t]),
    "lamcg_generate_rhs": (ctypes.c_int, [_vp]),
    "lamcg_load_matrix": (ctypes.c_int, [_vp, _cp]),
    "lamcg_load_rhs": (ctypes.c_int, [_vp, _cp]),
    "lamcg_set_matrix": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, ctypes.c_int]),
    "lamcg_set_rhs": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t]),
    "lamcg_random_spd_system": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_int]),
    "lamcg_save_system": (ctypes.c_int, [_vp, _cp, _cp]),
    "lamcg_solve": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, ctypes.POINTER(lamcg_result)]),
    "lamcg_get_residual_history": (ctypes.c_int, [_vp, _dp, ctypes.c_int]),
    "lamcg_get_solution_local": (ctypes.c_int, [_vp, _vp]),
    "lamcg_get_solution": (ctypes.c_int, [_vp, _vp]),
    "lamcg_save_solution": (ctypes.c_int, [_vp, _cp]),
    "lamcg_solve_resume": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, ctypes.POINTER(lamcg_result)]),
    "lamcg_checkpoint_save": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "lamcg_checkpoint_load": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "lamcg_gemv": (ctypes.c_int, [_vp, _vp, _vp, _dp]),
    "lamcg_time_gemv": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _dp]),
    "lamcg_vector_update_step": (ctypes.c_int, [_vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp, _dp, _dp]),
    "lamcg_get_loop_profile": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]),
    "lamcg_time_stream_read": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _dp, _dp]),
}


def lib() -> ctypes.CDLL:
    """Load liblamcg.so (fails loudly if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LamcgError(-2, f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(the CUDA library is the only implementation; there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _as_array(a, name: str, dtype=np.float64) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=dtype)
    if arr.size == 0:
        raise ValueError(f"{name} is empty")
    return arr


DTYPES = {"f64": (0, np.float64), "f32": (1, np.float32), np.float64: (0, np.float64), np.float32: (1, np.float32)}


class Solver:
    """Direct, exception-raising wrapper over one ``lamcg_t`` (one rank == one GPU)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, dtype="f64"):
        self._L = lib()
        h = _vp()
        self.dtype_code, self.np_dtype = DTYPES[dtype]
        rc = self._L.lamcg_create_typed(ctypes.byref(h), device, rank, nranks, self.dtype_code)
        if rc != 0:
            raise LamcgError(rc, (self._L.lamcg_last_error(None) or b"").decode())
        self._h = h
        self.rank, self.nranks, self.device = rank, nranks, device

    # -- plumbing
    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise LamcgError(rc, (self._L.lamcg_last_error(self._h) or b"").decode())
        return rc

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.lamcg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key: str, value: int) -> None:
        self._ck(self._L.lamcg_set_option(self._h, key.encode(), int(value)))

    @property
    def info(self) -> lamcg_info:
        out = lamcg_info()
        self._ck(self._L.lamcg_get_info(self._h, ctypes.byref(out)))
        return out

    # -- comm bootstrap
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
        rc = lib().lamcg_comm_nccl_unique_id(ctypes.cast(buf, _vp))
        if rc != 0:
            raise LamcgError(rc, (lib().lamcg_last_error(None) or b"").decode())
        return buf.raw

    def comm_init_nccl(self, unique_id: bytes) -> None:
        assert len(unique_id) == NCCL_ID_BYTES
        buf = ctypes.create_string_buffer(unique_id, NCCL_ID_BYTES)
        self._ck(self._L.lamcg_comm_init_nccl(self._h, ctypes.cast(buf, _vp)))

    def comm_peer_export(self, n: int) -> bytes:
        buf = ctypes.create_string_buffer(PEER_HANDLE_BYTES)
        self._ck(self._L.lamcg_comm_peer_export(self._h, n, ctypes.cast(buf, _vp)))
        return buf.raw

    def comm_init_peer(self, all_handles: bytes) -> None:
        assert len(all_handles) == PEER_HANDLE_BYTES * self.nranks
        buf = ctypes.create_string_buffer(all_handles, len(all_handles))
        self._ck(self._L.lamcg_comm_init_peer(self._h, ctypes.cast(buf, _vp)))

    # -- system
    def generate_matrix(self, rows: int, cols: int) -> None:
        self._ck(self._L.lamcg_generate_matrix(self._h, rows, cols))

    def generate_rhs(self) -> None:
        self._ck(self._L.lamcg_generate_rhs(self._h))

    def load_matrix(self, path: str) -> None:
        self._ck(self._L.lamcg_load_matrix(self._h, os.fsencode(path)))

    def load_rhs(self, path: str) -> None:
        self._ck(self._L.lamcg_load_rhs(self._h, os.fsencode(path)))

    def _check_matrix_shape(self, shape, layout: int) -> int:
        """layout 0: the whole (n, n) matrix; layout 1: this rank's (local_rows, n) block.  Anything else would make the
        library's 2-D copy read past the end of the caller's buffer, so it is refused here (ValueError) before calling into C."""
        if layout not in (0, 1):
            raise ValueError(f"layout must be 0 (whole matrix) or 1 (this rank's row block), not {layout!r}")
        if len(shape) != 2:
            raise ValueError(f"A must be 2-dimensional, got shape {tuple(shape)}")
        rows, n = int(shape[0]), int(shape[1])
        want = n if layout == 0 else launch.partition(n, self.nranks, self.rank)[0]
        if rows != want or n == 0:
            raise ValueError(f"A has shape {tuple(shape)}; layout {layout} on rank {self.rank}/{self.nranks} needs ({want}, {n})")
        return n

    def set_matrix(self, A, layout: int = 0) -> None:
        """A: numpy array (host) or an object with ``data_ptr()`` (torch tensor, host or device)."""
        if hasattr(A, "data_ptr"):
            if not A.is_contiguous() or A.element_size() != np.dtype(self.np_dtype).itemsize:
                raise ValueError("A must be contiguous and of the handle's element type")
            n = self._check_matrix_shape(tuple(A.shape), layout)
            self._keep_A = A
            self._ck(self._L.lamcg_set_matrix(self._h, _vp(A.data_ptr()), n, layout))
        else:
            arr = _as_array(A, "A", self.np_dtype)
            n = self._check_matrix_shape(arr.shape, layout)
            self._ck(self._L.lamcg_set_matrix(self._h, _vp(arr.ctypes.data), n, layout))

    def set_rhs(self, b) -> None:
        if hasattr(b, "data_ptr"):
            if not b.is_contiguous() or b.element_size() != np.dtype(self.np_dtype).itemsize:
                raise ValueError("b must be contiguous and of the handle's element type")
            self._ck(self._L.lamcg_set_rhs(self._h, _vp(b.data_ptr()), b.numel()))
        else:
            arr = _as_array(b, "b", self.np_dtype).reshape(-1)
            self._ck(self._L.lamcg_set_rhs(self._h, _vp(arr.ctypes.data), arr.size))

    def random_spd_system(self, n: int, seed: int) -> None:
        """GPU version of the reference's random_spd_system tool (fills A and b of this handle)."""
        self._ck(self._L.lamcg_random_spd_system(self._h, n, seed))

    def save_system(self, matrix_path: str, rhs_path: str) -> None:
        self._ck(self._L.lamcg_save_system(self._h, os.fsencode(matrix_path), os.fsencode(rhs_path)))

    # -- solve
    def solve(self, max_iters: int, rel_error: float) -> lamcg_result:
        out = lamcg_result()
        self._ck(self._L.lamcg_solve(self._h, int(max_iters), float(rel_error), ctypes.byref(out)))
        return out

    def solve_resume(self, more_iters: int, rel_error: float) -> lamcg_result:
        """Carry on the last solve (it must have stopped on max_iters) for up to ``more_iters`` iterations;
        bit-identical to one uninterrupted solve.  Totals are reported."""
        out = lamcg_result()
        self._ck(self._L.lamcg_solve_resume(self._h, int(more_iters), float(rel_error), ctypes.byref(out)))
        return out

    def checkpoint_save(self, path: str) -> None:
        """This rank's x, r, p slices + scalars + iteration count (one file per rank)."""
        self._ck(self._L.lamcg_checkpoint_save(self._h, os.fsencode(path)))

    def checkpoint_load(self, path: str) -> None:
        self._ck(self._L.lamcg_checkpoint_load(self._h, os.fsencode(path)))

    def residual_history(self, capacity: int | None = None) -> np.ndarray:
        cap = capacity if capacity is not None else 1 << 22
        buf = np.zeros(cap)
        cnt = self._ck(self._L.lamcg_get_residual_history(self._h, buf.ctypes.data_as(_dp), cap))
        return buf[:cnt].copy()

    def solution_local(self) -> np.ndarray:
        x = np.zeros(max(self.info.local_rows, 1), dtype=self.np_dtype)
        self._ck(self._L.lamcg_get_solution_local(self._h, _vp(x.ctypes.data)))
        return x[: self.info.local_rows]

    def solution(self, out=None) -> np.ndarray:
        """Whole x; ``out`` may be a pinned torch tensor / numpy array of n elements of the handle's type (host memory)."""
        n = self.info.n
        if out is None:
            out = np.zeros(n, dtype=self.np_dtype)
        if hasattr(out, "data_ptr"):
            size, item, contig = out.numel(), out.element_size(), out.is_contiguous()
        else:
            size, item, contig = out.size, out.itemsize, out.flags["C_CONTIGUOUS"]
        if size < n or item != np.dtype(self.np_dtype).itemsize or not contig:
            raise ValueError(f"out must be a contiguous buffer of at least n = {n} elements of {np.dtype(self.np_dtype).name}")
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._ck(self._L.lamcg_get_solution(self._h, _vp(ptr)))
        return out

    def save_solution(self, path: str) -> None:
        self._ck(self._L.lamcg_save_solution(self._h, os.fsencode(path)))

    # -- hooks
    def gemv(self, p) -> tuple[np.ndarray, float]:
        p = _as_array(p, "p", self.np_dtype).reshape(-1)
        y = np.zeros(max(self.info.local_rows, 1), dtype=self.np_dtype)
        d = ctypes.c_double()
        self._ck(self._L.lamcg_gemv(self._h, _vp(p.ctypes.data), _vp(y.ctypes.data), ctypes.byref(d)))
        return y[: self.info.local_rows], d.value

    def vector_update_step(self, x, r, p, Ap, rr: float, pAp: float, fused: bool = True):
        """K2 / K3 in isolation (test hook): returns (x_new, r_new, p_new, alpha, rr_new, beta)."""
        x, r, p = (np.array(v, dtype=self.np_dtype).reshape(-1) for v in (x, r, p))
        Ap = _as_array(Ap, "Ap", self.np_dtype).reshape(-1)
        a, rn, b = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._ck(self._L.lamcg_vector_update_step(self._h, x.size, _vp(x.ctypes.data), _vp(r.ctypes.data), _vp(p.ctypes.data), _vp(Ap.ctypes.data),
                                                  float(rr), float(pAp), int(fused), ctypes.byref(a), ctypes.byref(rn), ctypes.byref(b)))
        return x, r, p, a.value, rn.value, b.value

    def time_gemv(self, warmup: int = 3, reps: int = 10) -> float:
        ms = ctypes.c_double()
        self._ck(self._L.lamcg_time_gemv(self._h, warmup, reps, ctypes.byref(ms)))
        return ms.value

    def loop_profile(self) -> list[int]:
        """SM cycles per phase of the last persistent-loop solve (CTA 0): p update, GEMV, row sums + p.Ap
        exchange, alpha broadcast, x/r update + r.r exchange, beta broadcast."""
        buf = (ctypes.c_longlong * 8)()
        cnt = self._ck(self._L.lamcg_get_loop_profile(self._h, buf, 8))
        return [int(buf[i]) for i in range(min(cnt, 6))]

    def time_stream_read(self, warmup: int = 2, reps: int = 5) -> tuple[float, float]:
        ms, cs = ctypes.c_double(), ctypes.c_double()
        self._ck(self._L.lamcg_time_stream_read(self._h, warmup, reps, ctypes.byref(ms), ctypes.byref(cs)))
        return ms.value, cs.value


class ConjugateGradient_B200:
    """Mirror of the reference solver classes: bool returns, messages on stderr, never raises for
    I/O or shape errors (OMP.hpp:98-118), ``solve`` returns False when not converged (OMP.hpp:86-90).

    Same public methods as LAM::ConjugateGradient_CPU_MPI_OMP<double> (MPI_OMP.hpp:22-35)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, verbose: bool = True, dtype="f64"):
        self._s = Solver(device, rank, nranks, dtype)
        self.verbose = verbose
        self.last_result: lamcg_result | None = None

    @property
    def solver(self) -> Solver:
        return self._s

    def _try(self, fn, *a) -> bool:
        try:
            fn(*a)
            return True
        except LamcgError as e:
            if self._s.rank == 0:
                print(e.message, file=sys.stderr)
            return False

    def load_matrix_from_file(self, filename: str) -> bool:
        return self._try(self._s.load_matrix, filename)

    def load_rhs_from_file(self, filename: str) -> bool:
        return self._try(self._s.load_rhs, filename)

    def save_result_to_file(self, filename: str) -> bool:
        return self._try(self._s.save_solution, filename)

    def generate_matrix(self, rows: int, cols: int) -> bool:
        return self._try(self._s.generate_matrix, rows, cols)

    def generate_rhs(self) -> bool:
        return self._try(self._s.generate_rhs)

    def get_num_rows(self) -> int:
        return int(self._s.info.local_rows)  # local rows, like MPI_OMP.hpp:34

    def get_num_cols(self) -> int:
        return int(self._s.info.n)

    def solve(self, max_iters: int, rel_error: float) -> bool:
        res = self._s.solve(max_iters, rel_error)
        self.last_result = res
        if self.verbose and self._s.rank == 0:
            if res.converged:
                print("Converged in %d iterations, relative error is %e" % (res.iterations, res.rel_residual))
            else:
                print("Did not converge in %d iterations, relative error is %e" % (max_iters, res.rel_residual))
        return bool(res.converged)

    def resume(self, more_iters: int, rel_error: float) -> bool:
        """Continue a solve that did not converge within max_iters (no reference equivalent: it restarts from x = 0)."""
        try:
            res = self._s.solve_resume(more_iters, rel_error)
        except LamcgError as e:
            if self._s.rank == 0:
                print(e.message, file=sys.stderr)
            return False
        self.last_result = res
        if self.verbose and self._s.rank == 0:
            if res.converged:
                print("Converged in %d iterations, relative error is %e" % (res.iterations, res.rel_residual))
            else:
                print("Did not converge in %d iterations, relative error is %e" % (res.iterations - 1, res.rel_residual))
        return bool(res.converged)

    def _rank_path(self, filename: str) -> str:
        return filename if self._s.nranks == 1 else "%s.rank%dof%d" % (filename, self._s.rank, self._s.nranks)

    def save_checkpoint_to_file(self, filename: str) -> bool:
        """One file per rank (``<filename>.rank<r>of<P>`` when P > 1)."""
        return self._try(self._s.checkpoint_save, self._rank_path(filename))

    def load_checkpoint_from_file(self, filename: str) -> bool:
        return self._try(self._s.checkpoint_load, self._rank_path(filename))

    def solve_system(self, A, b, x, max_iters: int, rel_error: float) -> bool:
        """The original challenge signature solve(A, b, x, size, max_iters, rel_error)
        (test/test_CG_CPU_OMP.cpp:76-79) on caller-owned buffers; x is filled in place."""
        self._s.set_matrix(A)
        self._s.set_rhs(b)
        ok = self.solve(max_iters, rel_error)
        self._s.solution(out=x)
        return ok

    def close(self) -> None:
        self._s.close()
