"""Host helpers -- mirror of ``qgs/functions/util.py`` (small per-call arrays, not on the device path)."""
import numpy as np


def reverse(a):
    """Reverse a 1D array (util.py:34-53)."""
    return np.ascontiguousarray(np.asarray(a)[::-1])


def normalize_matrix_columns(a):
    """Normalise the columns of a 2D array; returns (normalised, norms) (util.py:56-75)."""
    a = np.asarray(a, dtype=np.float64)
    norm = np.zeros(a.shape[0])
    nrm = np.linalg.norm(a, 2, axis=0)
    norm[:a.shape[1]] = nrm
    return a / nrm[np.newaxis, :], norm


def solve_triangular_matrix(a, b):
    """Solve the triangular system column by column as util.py:78-98 does: column ``i-1`` of the solution uses the
    leading ``i x i`` block."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros_like(a)
    for i in range(2, a.shape[0] + 1):
        x[:i, i - 1] = np.linalg.solve(a[:i, :i], b[:i, i - 1])
    x[0, 0] = b[0, 0] / a[0, 0]
    return x
