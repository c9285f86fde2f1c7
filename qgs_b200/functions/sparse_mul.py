"""Sparse tensor contractions on the GPU -- mirror of ``qgs/functions/sparse_mul.py``.

Same names, argument order and return shapes as the reference's numba functions; the work is done
by ``qgsb_sparse_mul2/3/4/5`` (``include/qgsb.h``).  These are the raw contractions with arbitrary
vectors; the integrators never call them (they use a device-resident tensor handle, see
``qgs_b200.functions.tendencies``).
"""
import ctypes

import numpy as np

from qgs_b200 import _lib


def _coo(coo, rank):
    coo = np.asarray(coo)
    if coo.ndim != 2 or coo.shape[1] != rank:
        raise ValueError("coo must have shape (n_elems, %d), got %s" % (rank, coo.shape))
    return _lib.i32(coo)


def sparse_mul2(coo, value, vec):
    """``A_ij = sum_k T_ijk a_k``  (sparse_mul.py:13-45) -> (len(vec), len(vec))."""
    coo, value, vec = _coo(coo, 3), _lib.f64(value), _lib.f64(vec)
    res = np.empty((len(vec), len(vec)))
    _lib.check(_lib.load().qgsb_sparse_mul2(len(value), coo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(value),
                                            len(vec), _lib.dptr(vec), _lib.dptr(res)))
    return res


def sparse_mul3(coo, value, vec_a, vec_b):
    """``v_i = sum_jk T_ijk a_j b_k``, ``v_0 = 1``  (sparse_mul.py:48-81)."""
    coo, value, vec_a, vec_b = _coo(coo, 3), _lib.f64(value), _lib.f64(vec_a), _lib.f64(vec_b)
    res = np.empty_like(vec_a)
    _lib.check(_lib.load().qgsb_sparse_mul3(len(value), coo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(value),
                                            len(vec_a), _lib.dptr(vec_a), _lib.dptr(vec_b), _lib.dptr(res)))
    return res


def sparse_mul4(coo, value, vec_a, vec_b, vec_c):
    """``A_ij = sum_klm T_ijklm a_k b_l c_m``  (sparse_mul.py:84-118)."""
    coo, value = _coo(coo, 5), _lib.f64(value)
    vec_a, vec_b, vec_c = _lib.f64(vec_a), _lib.f64(vec_b), _lib.f64(vec_c)
    res = np.empty((len(vec_a), len(vec_a)))
    _lib.check(_lib.load().qgsb_sparse_mul4(len(value), coo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(value),
                                            len(vec_a), _lib.dptr(vec_a), _lib.dptr(vec_b), _lib.dptr(vec_c),
                                            _lib.dptr(res)))
    return res


def sparse_mul5(coo, value, vec_a, vec_b, vec_c, vec_d):
    """``v_i = sum_jklm T_ijklm a_j b_k c_l d_m``, ``v_0 = 1``  (sparse_mul.py:121-158)."""
    coo, value = _coo(coo, 5), _lib.f64(value)
    vec_a, vec_b, vec_c, vec_d = _lib.f64(vec_a), _lib.f64(vec_b), _lib.f64(vec_c), _lib.f64(vec_d)
    res = np.empty_like(vec_a)
    _lib.check(_lib.load().qgsb_sparse_mul5(len(value), coo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(value),
                                            len(vec_a), _lib.dptr(vec_a), _lib.dptr(vec_b), _lib.dptr(vec_c),
                                            _lib.dptr(vec_d), _lib.dptr(res)))
    return res
