"""Minimal coordinate-based ``sparse`` (pydata) look-alike: the subset used by qgs tensor construction.

Call sites in the reference: ``qgs/tensors/qgtensor.py:188-272, 657-746, 969-1005``,
``qgs/tensors/atmo_thermo_tensor.py:100-344`` and ``qgs/inner_products/analytic.py:131-216`` /
``symbolic.py:257-1457``:

* ``zeros(shape, dtype, format='dok' | 'coo')``; item get / set / ``+=`` / ``-=`` with integer tuples on the DOK form;
  ``to_coo()``;
* ``COO(dense)``, ``COO(coords, data, shape=, prune=)`` (duplicates summed);
* ``.coords`` ``(rank, nnz)``, ``.data``, ``.nnz``, ``.shape``, ``.copy()``, ``swapaxes``, ``+``, scalar ``*``, unary ``-``;
* indexing with integers and unit-step slices (``a_inv[i, :]``, ``aips._g[offset:, jo, offset:]``, ``a_theta[i]``,
  ``bips._Z[jj, j, k, ell, m]``);
* ``@`` between vectors and matrices (sparse or dense), ``tensordot(a, b, axes=1)``.

Storage is a real coordinate list -- an ``(rank, nnz)`` int64 array and an ``(nnz,)`` float64 array kept in lexicographic
order with duplicates summed and exact zeros dropped -- so memory is proportional to the number of non-zeros: the rank-5
tensor of the 38-variable T4 model is 5 340 entries (256 KB) instead of the 39**5 doubles (0.7 GB) a dense backing array
would need, and a rank-5 tensor at the 228 variables of the 6x6 model (229**5 doubles dense) becomes possible at all.
The lexicographic entry order is the order pydata ``sparse`` produces and therefore the summation order of the
reference's ``sparse_mul*`` loops.  Exact zeros are never stored: ``coords`` / ``data`` list non-zeros only.
"""
import numpy as np

__all__ = ["COO", "DOK", "zeros", "tensordot"]

# Arrays up to this many elements (32 MB dense) are sliced and multiplied through a dense copy built once: numpy's
# BLAS-backed products then give what qgs gets from its small dense inverse matrices (and what the tensor fixtures under
# tests/golden were built with), and the hundreds of thousands of slices qgs takes of its inner-product arrays cost
# O(slice).  Larger arrays -- the model tensors themselves, (ndim + 1)**3 or **5 elements -- only ever exist as
# coordinate lists.
DENSE_LIMIT = 1 << 22


def _canonical(coords, data, shape):
    """Lexicographic order, duplicates summed (in input order, like ``np.add.at``), zeros dropped."""
    coords = np.asarray(coords, dtype=np.int64).reshape(len(shape), -1)
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    if coords.shape[1] != data.shape[0]:
        raise ValueError("coords describe %d entries, data holds %d" % (coords.shape[1], data.shape[0]))
    if data.size == 0:
        return np.zeros((len(shape), 0), dtype=np.int64), np.zeros(0)
    for axis, extent in enumerate(shape):
        if coords[axis].min() < 0 or coords[axis].max() >= extent:
            raise IndexError("index out of bounds for axis %d with size %d" % (axis, extent))
    if len(shape) == 0:
        total = data.sum()
        return np.zeros((0, 1 if total != 0. else 0), dtype=np.int64), (np.array([total]) if total != 0. else np.zeros(0))
    flat = np.ravel_multi_index(tuple(coords), shape) if int(np.prod(shape, dtype=object)) < 2 ** 62 else None
    if flat is None:        # huge index spaces: sort on the coordinate columns themselves
        order = np.lexsort(coords[::-1])
        coords, data = coords[:, order], data[order]
        new = np.concatenate(([True], np.any(coords[:, 1:] != coords[:, :-1], axis=0)))
    else:
        order = np.argsort(flat, kind="stable")
        flat, coords, data = flat[order], coords[:, order], data[order]
        new = np.concatenate(([True], flat[1:] != flat[:-1]))
    if not new.all():
        # strictly sequential sums in input order (what np.add.at into a dense array does, and what the fixtures under
        # tests/golden were built with): np.add.reduceat would sum long runs pairwise and differ in the last bit
        starts = np.flatnonzero(new)
        summed = np.zeros(len(starts))
        np.add.at(summed, np.cumsum(new) - 1, data)
        data = summed
        coords = coords[:, starts]
    keep = data != 0.
    if not keep.all():
        coords, data = coords[:, keep], data[keep]
    return np.ascontiguousarray(coords), np.ascontiguousarray(data)


def _as_coo(x):
    if isinstance(x, COO):
        return x
    if isinstance(x, DOK):
        return x.to_coo()
    return COO(np.asarray(x, dtype=np.float64))


class COO(object):
    """Immutable N-dimensional sparse array in coordinate format."""

    __array_priority__ = 20.0          # numpy defers binary operators to this class
    __array_ufunc__ = None

    def __init__(self, coords, data=None, shape=None, prune=False, **_ignored):
        if data is None:
            if isinstance(coords, (COO, DOK)):
                other = _as_coo(coords)
                self.shape, self.coords, self.data = other.shape, other.coords.copy(), other.data.copy()
                return
            dense = np.asarray(coords, dtype=np.float64)
            self.shape = tuple(int(s) for s in dense.shape)
            nz = np.nonzero(dense)
            self.coords = np.array(nz, dtype=np.int64).reshape(dense.ndim, -1)
            self.data = np.ascontiguousarray(dense[nz], dtype=np.float64).reshape(-1)
            return
        coords = np.asarray(coords)
        if coords.ndim == 1:
            coords = coords[np.newaxis, :]
        if shape is None:
            shape = tuple(int(m) + 1 for m in coords.max(axis=1)) if coords.shape[1] else (0,) * coords.shape[0]
        self.shape = tuple(int(s) for s in shape)
        self.coords, self.data = _canonical(coords, data, self.shape)

    @classmethod
    def _raw(cls, coords, data, shape):
        """Build from arrays that are already canonical."""
        out = cls.__new__(cls)
        out.coords, out.data, out.shape = coords, data, tuple(int(s) for s in shape)
        return out

    # ---- attributes -----------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nnz(self):
        return int(self.data.shape[0])

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=object))

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def fill_value(self):
        return 0.0

    @property
    def T(self):
        return COO(self.coords[::-1], self.data, shape=self.shape[::-1])

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return "<COO: shape=%s, nnz=%d>" % (self.shape, self.nnz)

    def todense(self):
        return self._dense().copy()

    def _dense(self):
        """Dense view of a SMALL array, built once (the object is immutable).  qgs slices its inner-product arrays
        hundreds of thousands of times (``aips._b[offset:, jo, ko]`` for every j, k): slicing a dense copy is O(slice),
        masking the coordinate list would be O(nnz) per call."""
        cached = getattr(self, "_dense_cache", None)
        if cached is None:
            cached = np.zeros(self.shape)
            if self.nnz:
                cached[tuple(self.coords)] = self.data
            if self.size <= DENSE_LIMIT:
                self._dense_cache = cached
        return cached

    def __array__(self, dtype=None, copy=None):
        return self.todense() if dtype is None else self.todense().astype(dtype)

    def to_coo(self):
        return self

    def asformat(self, fmt):
        return DOK.from_coo(self) if fmt == "dok" else self

    def copy(self, order="C"):
        return COO._raw(self.coords.copy(), self.data.copy(), self.shape)

    def astype(self, dtype):
        return self.copy()

    # ---- structure ----------------------------------------------------------------------------------------
    def swapaxes(self, a, b):
        coords = self.coords.copy()
        coords[[a, b]] = coords[[b, a]]
        shape = list(self.shape)
        shape[a], shape[b] = shape[b], shape[a]
        return COO(coords, self.data, shape=tuple(shape))

    def transpose(self, axes=None):
        axes = tuple(range(self.ndim))[::-1] if axes is None else tuple(axes)
        return COO(self.coords[list(axes)], self.data, shape=tuple(self.shape[a] for a in axes))

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            at = [i for i, k in enumerate(key) if k is Ellipsis][0]
            key = key[:at] + (slice(None),) * (self.ndim - len(key) + 1) + key[at + 1:]
        if len(key) > self.ndim:
            raise IndexError("too many indices for a %d-dimensional array" % self.ndim)
        key = key + (slice(None),) * (self.ndim - len(key))
        if self.size <= DENSE_LIMIT:
            for k in key:
                if isinstance(k, slice) and k.step not in (None, 1):
                    raise NotImplementedError("only unit-step slices are supported")
            part = self._dense()[key]
            return COO(part) if isinstance(part, np.ndarray) and part.ndim > 0 else float(part)
        mask = np.ones(self.nnz, dtype=bool)
        kept, shape, shift = [], [], []
        for axis, k in enumerate(key):
            extent = self.shape[axis]
            if isinstance(k, (int, np.integer)):
                k = int(k)
                if k < 0:
                    k += extent
                if not 0 <= k < extent:
                    raise IndexError("index %d is out of bounds for axis %d with size %d" % (k, axis, extent))
                mask &= self.coords[axis] == k
            elif isinstance(k, slice):
                start, stop, step = k.indices(extent)
                if step != 1:
                    raise NotImplementedError("only unit-step slices are supported")
                stop = max(stop, start)
                mask &= (self.coords[axis] >= start) & (self.coords[axis] < stop)
                kept.append(axis)
                shape.append(stop - start)
                shift.append(start)
            else:
                raise NotImplementedError("index of type %s" % type(k).__name__)
        if not kept:
            return float(self.data[mask].sum())          # at most one entry: the array is canonical
        coords = self.coords[kept][:, mask] - np.asarray(shift, dtype=np.int64)[:, None]
        # a sub-selection of a lexicographically ordered list with some axes dropped is still ordered
        return COO._raw(np.ascontiguousarray(coords), np.ascontiguousarray(self.data[mask]), tuple(shape))

    # ---- arithmetic -----------------------------------------------------------------------------------------
    def __neg__(self):
        return COO._raw(self.coords, -self.data, self.shape)

    def __pos__(self):
        return self

    def __mul__(self, other):
        if isinstance(other, (COO, DOK, np.ndarray)) and np.ndim(other) > 0:
            other = _as_coo(other)
            if other.shape != self.shape:
                raise ValueError("shapes %s and %s differ" % (self.shape, other.shape))
            return COO(self.todense() * other.todense()) if self.size < (1 << 22) else NotImplemented
        return COO(self.coords, self.data * float(other), shape=self.shape)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return COO(self.coords, self.data / float(other), shape=self.shape)

    def __add__(self, other):
        if isinstance(other, (int, float, np.floating, np.integer)):
            if float(other) == 0.:
                return self
            raise ValueError("adding a non-zero scalar would make the array dense")
        other = _as_coo(other)
        if other.shape != self.shape:
            raise ValueError("shapes %s and %s differ" % (self.shape, other.shape))
        return COO(np.concatenate((self.coords, other.coords), axis=1), np.concatenate((self.data, other.data)),
                   shape=self.shape)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-_as_coo(other))

    def __rsub__(self, other):
        return (-self) + other

    def __matmul__(self, other):
        return _matmul(self, other)

    def __rmatmul__(self, other):
        return _matmul(other, self)

    def dot(self, other):
        return _matmul(self, other)

    def sum(self, axis=None):
        if axis is None:
            return float(self.data.sum())
        axes = (axis,) if isinstance(axis, (int, np.integer)) else tuple(axis)
        keep = [a for a in range(self.ndim) if a not in axes]
        return COO(self.coords[keep], self.data, shape=tuple(self.shape[a] for a in keep))

    def __eq__(self, other):
        if not isinstance(other, (COO, DOK)):
            return NotImplemented
        other = _as_coo(other)
        return self.shape == other.shape and np.array_equal(self.coords, other.coords) and \
            np.array_equal(self.data, other.data)

    __hash__ = None

    def __reduce__(self):
        return (_rebuild_coo, (self.coords, self.data, self.shape))


def _rebuild_coo(coords, data, shape):
    return COO._raw(coords, data, shape)


def _contract(a, b, axis_a, axis_b):
    """sum_k a[..., k, ...] b[..., k, ...] over one axis of each; the remaining axes of a come first."""
    a, b = _as_coo(a), _as_coo(b)
    if a.shape[axis_a] != b.shape[axis_b]:
        raise ValueError("shapes %s and %s are not aligned" % (a.shape, b.shape))
    rest_a = [x for x in range(a.ndim) if x != axis_a]
    rest_b = [x for x in range(b.ndim) if x != axis_b]
    shape = tuple(a.shape[x] for x in rest_a) + tuple(b.shape[x] for x in rest_b)
    if a.nnz == 0 or b.nnz == 0:
        return 0.0 if not shape else COO._raw(np.zeros((len(shape), 0), dtype=np.int64), np.zeros(0), shape)
    # join the two entry lists on the contracted index (entries of b grouped by it)
    order_b = np.argsort(b.coords[axis_b], kind="stable")
    kb = b.coords[axis_b][order_b]
    lo = np.searchsorted(kb, a.coords[axis_a], side="left")
    hi = np.searchsorted(kb, a.coords[axis_a], side="right")
    counts = hi - lo
    total = int(counts.sum())
    if total == 0:
        return 0.0 if not shape else COO._raw(np.zeros((len(shape), 0), dtype=np.int64), np.zeros(0), shape)
    ia = np.repeat(np.arange(a.nnz), counts)
    offs = np.arange(total) - np.repeat(np.cumsum(counts) - counts, counts)
    ib = order_b[np.repeat(lo, counts) + offs]
    data = a.data[ia] * b.data[ib]
    if not shape:
        return float(data.sum())
    coords = np.concatenate((a.coords[rest_a][:, ia], b.coords[rest_b][:, ib]), axis=0)
    return COO(coords, data, shape=shape)


def _wrap(result):
    if isinstance(result, np.ndarray) and result.ndim > 0:
        return COO(result)
    return float(result)


def _matmul(a, b):
    a, b = _as_coo(a), _as_coo(b)
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: scalar operands are not allowed")
    if a.size <= DENSE_LIMIT and b.size <= DENSE_LIMIT and a.ndim <= 2 and b.ndim <= 2:
        return _wrap(np.matmul(a._dense(), b._dense()))
    return _contract(a, b, a.ndim - 1, 0 if b.ndim == 1 else b.ndim - 2)


def tensordot(a, b, axes=2):
    """``axes=1``: last axis of ``a`` against the first of ``b`` (the only form qgs uses); ``axes=(i, j)``: one axis
    of each."""
    if isinstance(axes, (int, np.integer)):
        if axes != 1:
            raise NotImplementedError("tensordot over %d axes" % axes)
        a, b = _as_coo(a), _as_coo(b)
        if a.size <= DENSE_LIMIT and b.size <= DENSE_LIMIT:
            return _wrap(np.tensordot(a._dense(), b._dense(), axes=1))
        return _contract(a, b, a.ndim - 1, 0)
    ax_a, ax_b = axes
    if not isinstance(ax_a, (int, np.integer)):
        if len(ax_a) != 1:
            raise NotImplementedError("tensordot over several axes")
        ax_a, ax_b = ax_a[0], ax_b[0]
    return _contract(a, b, int(ax_a), int(ax_b))


class DOK(object):
    """Dictionary-of-keys form: mutable by item assignment with integer index tuples."""

    __array_priority__ = 20.0
    __array_ufunc__ = None

    def __init__(self, shape, dtype=np.float64, **_ignored):
        if np.isscalar(shape):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.entries = {}

    @classmethod
    def from_coo(cls, coo):
        out = cls(coo.shape)
        for idx, v in zip(zip(*coo.coords.tolist()), coo.data.tolist()):
            out.entries[idx] = v
        return out

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nnz(self):
        return sum(1 for v in self.entries.values() if v != 0.)

    def _key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) != self.ndim or not all(isinstance(k, (int, np.integer)) for k in key):
            return None
        out = []
        for axis, k in enumerate(key):
            k = int(k)
            if k < 0:
                k += self.shape[axis]
            if not 0 <= k < self.shape[axis]:
                raise IndexError("index %d is out of bounds for axis %d with size %d" % (k, axis, self.shape[axis]))
            out.append(k)
        return tuple(out)

    def __getitem__(self, key):
        full = self._key(key)
        if full is None:
            return self.to_coo()[key]
        return self.entries.get(full, 0.0)

    def __setitem__(self, key, value):
        full = self._key(key)
        if full is None:
            raise NotImplementedError("DOK assignment needs one integer per axis")
        if isinstance(value, (COO, DOK)):
            value = _as_coo(value)
            value = float(value.data.sum()) if value.nnz else 0.0
        value = float(np.asarray(value, dtype=np.float64).reshape(-1)[0]) if np.ndim(value) else float(value)
        if value == 0.:
            self.entries.pop(full, None)
        else:
            self.entries[full] = value

    def to_coo(self):
        if not self.entries:
            return COO._raw(np.zeros((self.ndim, 0), dtype=np.int64), np.zeros(0), self.shape)
        keys = np.array(list(self.entries.keys()), dtype=np.int64).reshape(-1, self.ndim).T
        return COO(keys, np.array(list(self.entries.values()), dtype=np.float64), shape=self.shape)

    def asformat(self, fmt):
        return self if fmt == "dok" else self.to_coo()

    def todense(self):
        return self.to_coo().todense()

    def __array__(self, dtype=None, copy=None):
        return self.todense()

    def copy(self):
        out = DOK(self.shape)
        out.entries = dict(self.entries)
        return out

    def __matmul__(self, other):
        return _matmul(self, other)

    def __rmatmul__(self, other):
        return _matmul(other, self)

    def __repr__(self):
        return "<DOK: shape=%s, nnz=%d>" % (self.shape, self.nnz)


def zeros(shape, dtype=np.float64, format="coo", **_ignored):
    if np.isscalar(shape):
        shape = (int(shape),)
    shape = tuple(int(s) for s in shape)
    if format == "dok":
        return DOK(shape)
    return COO._raw(np.zeros((len(shape), 0), dtype=np.int64), np.zeros(0), shape)
