"""numpy restatement of the reference's binary file format — TEST INFRASTRUCTURE ONLY.

Format (ref: challenge/main/random_spd_system.cpp:105-121 writer; LAM/src/CPU/
ConjugateGradient_CPU_OMP.hpp:137-197 matrix reader, :93-135 rhs reader, :199-217 solution
writer): 16-byte header = two native-endian ``size_t`` (rows, cols), then rows*cols float64,
row-major.  rhs / solution files have cols == 1.

Reference defect kept in mind (not reproduced): the solution writer stores an ``int num_cols=1``
with ``sizeof(size_t)`` (OMP.hpp:208-210), so bits 32..63 of the cols word are stack garbage in
files written by the reference; readers must mask with 0xffffffff.
"""
from __future__ import annotations

import numpy as np


def write_matrix(path: str, M: np.ndarray) -> None:
    M = np.ascontiguousarray(M, dtype=np.float64)
    if M.ndim == 1:
        M = M.reshape(-1, 1)
    with open(path, "wb") as f:
        np.array(M.shape, dtype=np.uint64).tofile(f)
        M.tofile(f)


def read_header(path: str) -> tuple[int, int]:
    hdr = np.fromfile(path, dtype=np.uint64, count=2)
    return int(hdr[0]), int(hdr[1])


def read_matrix(path: str, mask_cols: bool = False) -> np.ndarray:
    rows, cols = read_header(path)
    if mask_cols:
        cols &= 0xFFFFFFFF
    data = np.fromfile(path, dtype=np.float64, offset=16, count=rows * cols)
    assert data.size == rows * cols, "truncated file"
    return data.reshape(rows, cols)


def read_vector(path: str) -> np.ndarray:
    """rhs / solution file; tolerates the reference's garbage upper header bits."""
    return read_matrix(path, mask_cols=True).reshape(-1)
