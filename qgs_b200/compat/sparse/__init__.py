"""Minimal ``sparse`` (pydata) look-alike: the subset used by qgs tensor construction.

Call sites in the reference: ``qgs/tensors/qgtensor.py:188-272, 657-746, 969-1005`` and
``qgs/inner_products/analytic.py:131-216`` (``COO(dense)``, ``COO(coords, data, shape=, prune=)``,
``zeros(shape, dtype, format=)``, item get/set on the DOK form, ``to_coo``, ``@``, ``+``,
``swapaxes``, ``tensordot(axes=1)``, ``.coords/.data/.nnz``).

Storage is a plain dense ``numpy.ndarray`` subclass: fine up to rank 5 at ndim 38 (39**5 doubles =
0.7 GB) which is the largest tensor the reference's configurations build.  ``.coords`` is
``np.nonzero`` in C order, i.e. lexicographic in (i, j, k, ...), which is the entry order pydata
``sparse`` also produces and therefore the summation order of the reference ``sparse_mul*`` loops.
"""
import numpy as np

__all__ = ["COO", "DOK", "zeros", "tensordot"]


class COO(np.ndarray):
    __array_priority__ = 20.0

    def __new__(cls, coords, data=None, shape=None, prune=False, **_ignored):
        if data is None:
            arr = np.array(coords, dtype=np.float64, copy=True) if not isinstance(coords, np.ndarray) \
                else np.asarray(coords).astype(np.float64, copy=True)
            return arr.view(cls)
        coords = np.asarray(coords)
        data = np.asarray(data, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[np.newaxis, :]
        if shape is None:
            shape = tuple(int(m) + 1 for m in coords.max(axis=1))
        dense = np.zeros(tuple(shape), dtype=np.float64)
        if data.size:
            np.add.at(dense, tuple(coords.astype(np.intp)), data)
        return dense.view(cls)

    # ---- COO attributes -------------------------------------------------------------------
    @property
    def coords(self):
        return np.array(np.nonzero(np.asarray(self)), dtype=np.int64)

    @property
    def data(self):
        a = np.asarray(self)
        return a[np.nonzero(a)]

    @property
    def nnz(self):
        return int(np.count_nonzero(np.asarray(self)))

    @property
    def fill_value(self):
        return 0.0

    def todense(self):
        return np.array(self)

    def to_coo(self):
        return self

    def asformat(self, _fmt):
        return self

    def copy(self, order="C"):
        return np.array(self, copy=True).view(type(self))

    def __reduce__(self):
        return (_rebuild, (np.asarray(self).copy(), type(self).__name__))

    def __array_finalize__(self, obj):
        pass

    def __getitem__(self, key):
        out = np.ndarray.__getitem__(self, key)
        if isinstance(out, np.ndarray) and out.ndim == 0:
            return out[()]
        return out

    def __matmul__(self, other):
        out = np.matmul(np.asarray(self), np.asarray(other))
        return out.view(COO) if isinstance(out, np.ndarray) and out.ndim > 0 else out

    def __rmatmul__(self, other):
        out = np.matmul(np.asarray(other), np.asarray(self))
        return out.view(COO) if isinstance(out, np.ndarray) and out.ndim > 0 else out


class DOK(COO):
    """Dictionary-of-keys flavour: same dense storage, mutable by item assignment."""


def _rebuild(arr, name):
    return arr.view(DOK if name == "DOK" else COO)


def zeros(shape, dtype=np.float64, format="coo", **_ignored):
    if np.isscalar(shape):
        shape = (int(shape),)
    cls = DOK if format == "dok" else COO
    return np.zeros(tuple(shape), dtype=np.float64).view(cls)


def tensordot(a, b, axes=2):
    out = np.tensordot(np.asarray(a), np.asarray(b), axes=axes)
    return out.view(COO) if isinstance(out, np.ndarray) and out.ndim > 0 else out
