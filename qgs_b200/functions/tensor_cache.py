"""On-disk hand-off format of the tendencies tensor.

The reference rebuilds its tensor on every ``create_tendencies`` call (tendencies.py:57-96); for the rank-5 T4
configurations that construction takes minutes while the arrays that reach the hot path are a few hundred KB.
The file written here holds exactly those arrays -- ``ndim``, ``coo (nnz, rank)``, ``val``, ``jcoo``, ``jval``, the
transposed ``coords`` / ``data`` of ``QgsTensor.tensor`` and ``.jacobian_tensor`` (tendencies.py:92-96) -- as an
uncompressed ``.npz``; it is the format of the fixtures under ``tests/golden/tensor_*.npz``.

With ``QGSB_TENSOR_CACHE=<directory>`` set, ``create_tendencies(params)`` stores the arrays under a fingerprint of
the whole parameter object and later calls with equal parameters skip the tensor construction.  The cache is
bypassed whenever the caller asks for the inner products or the tensor object itself, which only the construction
can provide.
"""
import hashlib
import os

import numpy as np

FORMAT_VERSION = 1


def save_tensor(path, ndim, coo, val, jcoo, jval):
    """Write the tensor arrays to ``path`` (atomically: a reader never sees a partial file)."""
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    jcoo = np.ascontiguousarray(jcoo, dtype=np.int32)
    tmp = "%s.%d.tmp.npz" % (path, os.getpid())
    np.savez(tmp, format_version=FORMAT_VERSION, ndim=int(ndim), coo=coo, val=np.asarray(val, dtype=np.float64),
             jcoo=jcoo, jval=np.asarray(jval, dtype=np.float64))
    os.replace(tmp, path)
    return path


def load_tensor(path):
    """Read a tensor file; returns ``(ndim, coo, val, jcoo, jval)``.  Validates shapes so that a truncated or foreign
    file raises here instead of producing a wrong model."""
    with np.load(path) as z:
        version = int(z["format_version"]) if "format_version" in z else FORMAT_VERSION   # the golden fixtures predate it
        if version != FORMAT_VERSION:
            raise ValueError("%s: tensor file format %d, this build reads format %d" % (path, version, FORMAT_VERSION))
        ndim = int(z["ndim"])
        coo, val, jcoo, jval = z["coo"], z["val"], z["jcoo"], z["jval"]
    if coo.ndim != 2 or coo.shape[1] not in (3, 5) or len(val) != coo.shape[0]:
        raise ValueError("%s: not a tendencies tensor file (coo %s, val %s)" % (path, coo.shape, val.shape))
    if jcoo.shape[0] and (jcoo.ndim != 2 or jcoo.shape[1] != coo.shape[1] or len(jval) != jcoo.shape[0]):
        raise ValueError("%s: Jacobian tensor arrays are inconsistent (jcoo %s, jval %s)" % (path, jcoo.shape, jval.shape))
    if coo.size and (coo.min() < 0 or coo.max() > ndim):
        raise ValueError("%s: indices outside [0, %d]" % (path, ndim))
    return ndim, coo, val, jcoo, jval


def _walk(obj, h, stack):
    """Feed a canonical description of ``obj`` into the hash ``h``: values, not identities -- no ``id``-based reprs."""
    if obj is None or isinstance(obj, (bool, int, float, complex, str, bytes)):
        h.update(("%s:%r;" % (type(obj).__name__, obj)).encode())
        attrs = getattr(obj, "__dict__", None)             # qgs Parameter: a float carrying units / scaling flags
        if attrs:
            _walk(attrs, h, stack)
        return
    if isinstance(obj, np.generic):
        h.update(("np:%s:%r;" % (obj.dtype.str, obj.item())).encode())
        return
    if isinstance(obj, np.ndarray):
        h.update(("nd:%s:%s;" % (obj.dtype.str, obj.shape)).encode())
        if obj.dtype == object:
            for item in obj.ravel():
                _walk(item, h, stack)
        else:
            h.update(np.ascontiguousarray(obj).tobytes())
        return
    if id(obj) in stack:
        h.update(b"cycle;")
        return
    stack = stack | {id(obj)}
    if isinstance(obj, dict):
        h.update(b"{")
        for key in sorted(obj, key=repr):
            _walk(key, h, stack)
            _walk(obj[key], h, stack)
        h.update(b"}")
        return
    if isinstance(obj, (list, tuple)):
        h.update(b"[")
        for item in obj:
            _walk(item, h, stack)
        h.update(b"]")
        return
    if isinstance(obj, (set, frozenset)):
        h.update(b"<")
        for item in sorted(obj, key=repr):
            _walk(item, h, stack)
        h.update(b">")
        return
    kind = type(obj)
    h.update(("%s.%s:" % (kind.__module__, kind.__qualname__)).encode())
    if hasattr(obj, "free_symbols") and hasattr(obj, "args"):     # sympy expression (symbolic basis functions)
        h.update(str(obj).encode())
        return
    if callable(obj) and hasattr(obj, "__qualname__"):
        h.update(("%s.%s;" % (getattr(obj, "__module__", ""), obj.__qualname__)).encode())
        return
    attrs = getattr(obj, "__dict__", None)
    if attrs is not None:
        _walk(attrs, h, stack)
    slots = getattr(kind, "__slots__", ())
    for name in ([slots] if isinstance(slots, str) else slots):
        if hasattr(obj, name):
            _walk(name, h, stack)
            _walk(getattr(obj, name), h, stack)


_BUILDER_KEY = None


def _builder_key():
    """Version of the reference package plus a digest of the source of the modules that BUILD the tensor
    (qgs/tensors, qgs/inner_products): equal parameters must not be served a tensor an older construction produced.
    Under the overlay ``import qgs`` is the overlay package; it re-exports the real ``__version__`` and its
    ``__path__`` leads to the real sub-packages."""
    global _BUILDER_KEY
    if _BUILDER_KEY is None:
        key = hashlib.sha256()
        try:
            import qgs
            key.update(str(getattr(qgs, "__version__", "")).encode())
            import glob
            for base in list(getattr(qgs, "__path__", [])):
                for sub in ("tensors", "inner_products"):
                    for src in sorted(glob.glob(os.path.join(base, sub, "*.py"))):
                        with open(src, "rb") as fh:
                            key.update(os.path.basename(src).encode() + b"\0" + fh.read())
        except (ImportError, OSError):
            pass
        _BUILDER_KEY = key.hexdigest()
    return _BUILDER_KEY


def fingerprint(params, *extra):
    """Hex digest that changes whenever any value reachable from ``params`` (scalars, arrays, mode blocks, symbolic basis
    functions, nested parameter containers) changes."""
    h = hashlib.sha256()
    h.update(b"qgsb-tensor-v%d;" % FORMAT_VERSION)
    h.update(_builder_key().encode())
    _walk(params, h, frozenset())
    for item in extra:
        _walk(item, h, frozenset())
    return h.hexdigest()


def cache_file(params, kind="tendencies"):
    """Path of the cache entry for ``params`` under ``$QGSB_TENSOR_CACHE``, or ``None`` when caching is off."""
    root = os.environ.get("QGSB_TENSOR_CACHE")
    if not root:
        return None
    os.makedirs(root, exist_ok=True)
    return os.path.join(root, "%s_%s.npz" % (kind, fingerprint(params, kind)[:32]))
