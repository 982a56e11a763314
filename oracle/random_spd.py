"""numpy restatement of the reference's random SPD system generator — TEST INFRASTRUCTURE ONLY.

ref: challenge/main/random_spd_system.cpp
  * random_matrix (:27-38): glibc ``srand(seed)``; column-major fill with ``2*rand()/RAND_MAX - 1``.
  * random_spd_matrix (:66-101): Q = Gram-Schmidt-orthonormalised random matrix (seed);
    D = exp(3.5 * U(-1,1)) drawn with ``seed - 10``; A = (Q sqrt(D)) (Q sqrt(D))^T, so
    cond(A) <= e^7 ~ 1097.
  * main (:166): rhs = random_matrix(n, 1, seed + 10).

Third-party arithmetic: the reference orthonormalises with Intel MKL ``cblas_dnrm2/dscal/dgemm``
(recursive block Gram-Schmidt, :41-62), un-vendored and version-unpinned
(``module load intel``, challenge/random_spd_system.sh:6).  MKL is absent here and no reference
test pins the generator's output, so file-mode INPUTS are "parity unpinned" at the bit level: this
restates the *distribution* (Gram-Schmidt == thin QR with a positive diagonal of R) with the very
same glibc random stream.  Parity of the solver is unaffected: both solvers read the same file.
"""
from __future__ import annotations

import numpy as np

from . import rand_fill


def random_matrix(num_rows: int, num_cols: int, seed: int) -> np.ndarray:
    """Column-major fill, returned as a (num_rows, num_cols) array M with M[r, c] = stream[c*num_rows + r]."""
    flat = rand_fill(num_rows * num_cols, seed)
    return np.ascontiguousarray(flat.reshape(num_cols, num_rows).T)


def random_spd_matrix(n: int, seed: int) -> np.ndarray:
    M = random_matrix(n, n, seed)
    Q, R = np.linalg.qr(M)
    Q = Q * np.sign(np.diag(R))  # Gram-Schmidt convention: positive diagonal of R
    D = np.exp(3.5 * random_matrix(n, 1, seed - 10).reshape(-1))
    QD = Q * np.sqrt(D)  # scale column c by sqrt(D[c])
    A = QD @ QD.T
    # The reference computes A column-major with dgemm; A is symmetric, so the row-major file holds the same matrix.
    return np.ascontiguousarray(A)


def random_spd_system(n: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    A = random_spd_matrix(n, seed)
    b = random_matrix(n, 1, seed + 10).reshape(-1).copy()
    return A, b
