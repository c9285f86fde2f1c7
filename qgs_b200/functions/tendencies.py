"""Tendencies on the GPU -- mirror of ``qgs/functions/tendencies.py``.

``create_tendencies(params)`` keeps the reference's signature and optional returns
(tendencies.py:20-130).  The model tensor is still *built* by the reference's own (unchanged, out of
scope) ``qgs.inner_products`` / ``qgs.tensors`` packages; what changes is what happens to the
``coo, val, jcoo, jval`` arrays extracted at tendencies.py:92-96: instead of being captured by numba
closures they are uploaded once to the device (``qgsb_tensor_create``), and the returned ``f`` and
``Df`` are callables ``(t, x)`` that carry that device handle so that the integrators can run fused
kernels without ever calling back into Python.

``tendencies_from_tensor`` builds the same pair directly from the four arrays (any polynomial
system of degree <= 2 (rank 3) or <= 4 (rank 5), e.g. Lorenz-63/84).
"""
import ctypes
import os

import numpy as np

from qgs_b200 import _lib


class DeviceTensor(object):
    """Device-resident copy of the tendencies tensor and its Jacobian tensor (``qgsb_tensor``)."""

    def __init__(self, ndim, coo, val, jcoo=None, jval=None, specialise=True):
        coo = np.asarray(coo)
        if coo.ndim != 2 or coo.shape[1] not in (3, 5):
            raise ValueError("coo must have shape (n_elems, 3) or (n_elems, 5), got %s" % (coo.shape,))
        self.ndim = int(ndim)
        self.rank = int(coo.shape[1])
        self.coo = _lib.i32(coo)            # tendencies.py:92 hands an F-ordered int64 view
        self.val = _lib.f64(val)
        if jcoo is None:
            jcoo = np.zeros((0, self.rank), dtype=np.int32)
            jval = np.zeros((0,))
        self.jcoo = _lib.i32(jcoo)
        self.jval = _lib.f64(jval)
        if len(self.val) != self.coo.shape[0] or len(self.jval) != self.jcoo.shape[0]:
            raise ValueError("coordinate and value arrays have different lengths")
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().qgsb_tensor_create(
            self.ndim, self.rank, len(self.val), self.coo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(self.val),
            len(self.jval), self.jcoo.ctypes.data_as(_lib.c_int32_p), _lib.dptr(self.jval), ctypes.byref(handle)))
        self._handle = handle
        if specialise and self.kernel_kind != 2:
            self.specialise()

    # -- handle management -------------------------------------------------------------------------
    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("tensor handle already released")
        return self._handle

    def close(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().qgsb_tensor_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self.close()

    def _info(self):
        ndim, rank, kind = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        nnz, jnnz = ctypes.c_long(), ctypes.c_long()
        h = ctypes.c_uint64()
        _lib.check(_lib.load().qgsb_tensor_info(self.handle, ctypes.byref(ndim), ctypes.byref(rank), ctypes.byref(nnz),
                                                ctypes.byref(jnnz), ctypes.byref(kind), ctypes.byref(h)))
        return kind.value, h.value

    @property
    def kernel_kind(self):
        """0 generic thread-per-member, 1 generic warp-per-member, 2 tensor-specialised."""
        return self._info()[0]

    @property
    def tensor_hash(self):
        return self._info()[1]

    def use_specialised(self, enable=True):
        _lib.check(_lib.load().qgsb_tensor_use_specialised(self.handle, 1 if enable else 0))

    def specialise(self):
        """Build (or fetch from the cache) the tensor-specialised Runge-Kutta / tendencies kernels with nvcc (a few
        seconds) and attach them.  Returns True when the handle now runs specialised kernels; silently stays on the
        generic CUDA kernels when the tensor is too large or nvcc is not installed."""
        from qgs_b200 import codegen
        path = codegen.build_plugin(self.ndim, self.rank, self.coo, self.val, part="rk")
        if path is None:
            return False
        _lib.check(_lib.load().qgsb_load_plugin(path.encode()))
        self.use_specialised(True)
        return self.kernel_kind == 2

    @property
    def has_tangent(self):
        """True when generated tangent-linear / Benettin kernels are attached to this handle."""
        return bool(_lib.load().qgsb_tensor_has_tangent(self.handle))

    # member-steps x tangent vectors above which building the tangent kernels (tens of seconds of nvcc) pays off
    TANGENT_JIT_WORK = 5e7

    def ensure_tangent(self, work=None):
        """Called by the tangent-linear integrator and the Lyapunov estimators before a launch of ``work`` =
        members x steps x vectors (``None``: unconditionally): a handle
        that runs run-time built specialised kernels gets the matching tangent kernels too (unrolled Householder QR,
        four kernels: tens of seconds of nvcc once per tensor, cached on disk next to the first part).  Handles of the
        prebuilt configurations already have them; without nvcc the generic tangent kernels serve."""
        if work is not None and work < self.TANGENT_JIT_WORK:
            return False
        if self.has_tangent or self.kernel_kind != 2 or getattr(self, "_tangent_tried", False):
            return self.has_tangent
        self._tangent_tried = True
        if os.environ.get("QGSB_JIT_TANGENT", "1") == "0":
            return False
        from qgs_b200 import codegen
        path = codegen.build_plugin(self.ndim, self.rank, self.coo, self.val, jcoo=self.jcoo, jval=self.jval,
                                    part="tangent")
        if path is None:
            return False
        _lib.check(_lib.load().qgsb_load_plugin(path.encode()))
        self.use_specialised(True)
        return self.has_tangent


class _TensorCallable(object):
    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def ndim(self):
        return self.tensor.ndim

    def _states(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim not in (1, 2) or x.shape[-1] != self.tensor.ndim:
            # the reference's closures fail inside numba for a wrong length; integrators rely on that
            # to infer the dimension (integrator.py:344-361)
            raise ValueError("state must have %d components, got shape %s" % (self.tensor.ndim, x.shape))
        return x, _lib.f64(np.atleast_2d(x))


class Tendencies(_TensorCallable):
    """``f(t, x)`` of tendencies.py:99-115.  ``x`` may be one state ``(n,)`` or an ensemble ``(N, n)``."""

    def __call__(self, t, x):
        x, x2 = self._states(x)
        out = np.empty_like(x2)
        _lib.check(_lib.load().qgsb_tendencies(self.tensor.handle, x2.shape[0], _lib.dptr(x2), _lib.dptr(out)))
        return out.reshape(x.shape)


class JacobianTendencies(_TensorCallable):
    """``Df(t, x)`` of tendencies.py:105-121: ``(n,) -> (n, n)``, ``(N, n) -> (N, n, n)``."""

    def __call__(self, t, x):
        x, x2 = self._states(x)
        n = self.tensor.ndim
        out = np.empty((x2.shape[0], n, n))
        _lib.check(_lib.load().qgsb_jacobian(self.tensor.handle, x2.shape[0], _lib.dptr(x2), _lib.dptr(out)))
        return out[0] if x.ndim == 1 else out


def tendencies_from_tensor(ndim, coo, val, jcoo=None, jval=None, specialise=True):
    """Return ``(f, Df)`` for the polynomial system defined by a COO tensor (and its Jacobian tensor).

    ``coo`` is ``(n_elems, 3)`` or ``(n_elems, 5)`` with index 0 standing for the constant 1, exactly
    the arrays ``create_tendencies`` extracts at tendencies.py:92-96.  Without ``jcoo`` the Jacobian
    tensor is derived like ``QgsTensor.jacobian_from_tensor`` (qgtensor.py:700-722).
    """
    coo = np.asarray(coo)
    val = np.asarray(val, dtype=np.float64)
    if jcoo is None:
        jcoo, jval = jacobian_tensor_from_coo(coo, val)
    tensor = DeviceTensor(ndim, coo, val, jcoo, jval, specialise=specialise)
    return Tendencies(tensor), JacobianTendencies(tensor)


def tendencies_from_file(path, specialise=True):
    """``(f, Df)`` from a tensor file written by ``save_tendencies`` / the ``QGSB_TENSOR_CACHE`` cache (row f-1: the
    arrays of tendencies.py:92-96 without re-running the tensor construction)."""
    from qgs_b200.functions import tensor_cache
    ndim, coo, val, jcoo, jval = tensor_cache.load_tensor(path)
    return tendencies_from_tensor(ndim, coo, val, jcoo if len(jval) else None, jval if len(jval) else None,
                                  specialise=specialise)


def save_tendencies(path, f):
    """Write the tensor carried by ``f`` (or ``Df``) to ``path``."""
    from qgs_b200.functions import tensor_cache
    t = f.tensor
    return tensor_cache.save_tensor(path, t.ndim, t.coo, t.val, t.jcoo, t.jval)


def jacobian_tensor_from_coo(coo, val):
    """``T + sum_p swapaxes(T, 1, p+1)`` on coordinate lists (qgtensor.py:714-722), duplicates summed,
    entries returned in lexicographic order like pydata ``sparse`` does."""
    coo = np.asarray(coo, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    rank = coo.shape[1]
    parts_c, parts_v = [coo], [val]
    for p in range(1, rank - 1):
        sw = coo.copy()
        sw[:, [1, p + 1]] = sw[:, [p + 1, 1]]
        parts_c.append(sw)
        parts_v.append(val)
    allc = np.concatenate(parts_c)
    allv = np.concatenate(parts_v)
    uniq, inv = np.unique(allc, axis=0, return_inverse=True)
    summed = np.zeros(len(uniq))
    np.add.at(summed, inv.ravel(), allv)
    keep = summed != 0.
    return uniq[keep], summed[keep]


def _build_reference_tensor(params, thermo=False):
    """Run the reference's (out-of-scope) tensor construction.  tendencies.py:57-90 / :155-186."""
    try:
        from qgs.inner_products.analytic import AtmosphericAnalyticInnerProducts, OceanicAnalyticInnerProducts, \
            GroundAnalyticInnerProducts
    except ImportError:
        from qgs_b200 import compat
        compat.install()
        try:
            from qgs.inner_products.analytic import AtmosphericAnalyticInnerProducts, \
                OceanicAnalyticInnerProducts, GroundAnalyticInnerProducts
        except ImportError as exc:
            raise ImportError("create_tendencies needs the qgs package (parameters, inner products and tensor "
                              "construction are used from the reference unchanged): %s.  Use "
                              "tendencies_from_tensor(ndim, coo, val, jcoo, jval) when the tensor arrays are "
                              "already available." % exc)
    from qgs.inner_products.symbolic import AtmosphericSymbolicInnerProducts, OceanicSymbolicInnerProducts, \
        GroundSymbolicInnerProducts

    if params.ablocks is not None:
        aip = AtmosphericAnalyticInnerProducts(params)
    elif params.atmospheric_basis is not None:
        aip = AtmosphericSymbolicInnerProducts(params)
    else:
        aip = None

    if params.oblocks is not None:
        oip = OceanicAnalyticInnerProducts(params)
    elif params.oceanic_basis is not None:
        oip = OceanicSymbolicInnerProducts(params)
    else:
        oip = None

    if params.gblocks is not None:
        gip = GroundAnalyticInnerProducts(params)
    elif params.ground_basis is not None:
        gip = GroundSymbolicInnerProducts(params)
    else:
        gip = None

    if aip is not None and oip is not None:
        if not aip.connected_to_ocean:
            aip.connect_to_ocean(oip)
    elif aip is not None and gip is not None:
        if not aip.connected_to_ground:
            aip.connect_to_ground(gip)

    if thermo:
        from qgs.tensors.atmo_thermo_tensor import AtmoThermoTensor, AtmoThermoTensorDynamicT, AtmoThermoTensorT4
        classes = (AtmoThermoTensorT4, AtmoThermoTensorDynamicT, AtmoThermoTensor)
    else:
        from qgs.tensors.qgtensor import QgsTensor, QgsTensorDynamicT, QgsTensorT4
        classes = (QgsTensorT4, QgsTensorDynamicT, QgsTensor)
    if params.T4:
        agotensor = classes[0](params, aip, oip, gip)
    elif params.dynamic_T:
        agotensor = classes[1](params, aip, oip, gip)
    else:
        agotensor = classes[2](params, aip, oip, gip)
    return aip, oip, gip, agotensor


def create_tendencies(params, return_inner_products=False, return_qgtensor=False):
    """Same contract as the reference (tendencies.py:20-130): returns ``[f, Df, (inner products)?, (qgtensor)?]``
    with ``f(t, x)`` the tendencies and ``Df(t, x)`` the Jacobian matrix, both evaluated on the GPU."""
    from qgs_b200.functions import tensor_cache
    cached = None if (return_inner_products or return_qgtensor) else tensor_cache.cache_file(params)
    if cached is not None and os.path.exists(cached):
        ndim, coo, val, jcoo, jval = tensor_cache.load_tensor(cached)
        if ndim != params.ndim:
            raise RuntimeError("%s holds a %d-variable tensor, the parameters describe %d variables"
                               % (cached, ndim, params.ndim))
        aip = oip = gip = agotensor = None
    else:
        aip, oip, gip, agotensor = _build_reference_tensor(params)

        coo = agotensor.tensor.coords.T
        val = agotensor.tensor.data
        jcoo = agotensor.jacobian_tensor.coords.T
        jval = agotensor.jacobian_tensor.data
        store = tensor_cache.cache_file(params)
        if store is not None and not os.path.exists(store):
            tensor_cache.save_tensor(store, params.ndim, coo, val, jcoo, jval)

    tensor = DeviceTensor(params.ndim, coo, val, jcoo, jval)
    f = Tendencies(tensor)
    Df = JacobianTendencies(tensor)

    ret = list()
    ret.append(f)
    ret.append(Df)
    if return_inner_products:
        ret.append((aip, oip, gip))
    if return_qgtensor:
        ret.append(agotensor)
    return ret


def create_atmo_thermo_tendencies(params, return_atmo_thermo_tensor=False):
    """Same contract as tendencies.py:133-211: the partial thermodynamic tendencies used by the vertical wind
    diagnostic, evaluated with the same GPU contraction."""
    from qgs_b200.functions import tensor_cache
    cached = None if return_atmo_thermo_tensor else tensor_cache.cache_file(params, "atmo_thermo")
    if cached is not None and os.path.exists(cached):
        ndim, coo, val, _, _ = tensor_cache.load_tensor(cached)
        if ndim != params.ndim:
            raise RuntimeError("%s holds a %d-variable tensor, the parameters describe %d variables"
                               % (cached, ndim, params.ndim))
        agotensor = None
    else:
        aip, oip, gip, agotensor = _build_reference_tensor(params, thermo=True)

        coo = agotensor.tensor.coords.T
        val = agotensor.tensor.data
        store = tensor_cache.cache_file(params, "atmo_thermo")
        if store is not None and not os.path.exists(store):
            tensor_cache.save_tensor(store, params.ndim, coo, val, np.zeros((0, np.shape(coo)[1]), dtype=np.int32),
                                     np.zeros(0))

    f = Tendencies(DeviceTensor(params.ndim, coo, val, specialise=False))

    if return_atmo_thermo_tensor:
        ret = list()
        ret.append(f)
        ret.append(agotensor)
    else:
        ret = f
    return ret
