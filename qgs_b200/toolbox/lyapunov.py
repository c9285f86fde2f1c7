"""Lyapunov vectors and exponents on the GPU -- mirror of ``qgs/toolbox/lyapunov.py``.

``LyapunovsEstimator`` (Benettin BLV / FLV, lyapunov.py:41-632) and ``CovariantLyapunovsEstimator``
(Ginelli / subspace intersection, lyapunov.py:635-1329) keep the reference's constructors, methods,
attributes and return conventions.  The per-member worker loops -- propagate the tangent basis with
micro-steps ``mdt`` over each step ``dt``, ``np.linalg.qr``, ``log|diag R| / dt`` -- run as one
CUDA launch for the whole ensemble (``qgsb_lyap_benettin``): the basis is propagated directly
(``prop @ q`` by linearity), re-orthonormalised by a Householder QR with LAPACK's sign convention,
and the nonlinear trajectory is recomputed on the fly instead of being stored (BLV) or kept in HBM
(FLV).  The backward passes of the CLV methods run on the device too: the Ginelli recursion in
``qgsb_clv_ginelli``, the subspace intersections in ``qgsb_clv_subspace_intersect``.

The random start basis of the reference comes from numba's unseeded generator (lyapunov.py:592-593).
``LyapunovsEstimator`` draws and factorises it ON THE DEVICE (``qgsb_lyap_benettin`` with ``q0 = NULL``: a
counter-based uniform generator keyed by a seed and the member's index, then the kernel's own Householder
QR), so neither host random numbers nor a stacked LAPACK QR of ``(n_traj, n_dim, n_vec)`` stand in front
of the launch; the seed itself comes from ``numpy.random``, so ``np.random.seed`` makes runs reproducible.
``benettin(..., q0=<array>)`` still takes a host basis (the golden tests feed the reference's own draws).
"""
import ctypes
import multiprocessing

import numpy as np

from qgs_b200 import _lib
from qgs_b200.functions.util import normalize_matrix_columns, solve_triangular_matrix
from qgs_b200.integrators.integrate import (_integrate_runge_kutta_jit, _integrate_runge_kutta_tgls_jit, _zeros_func,
                                            n_records_of, rk4_tableau, tensor_of)


_SUBTIMES_CACHE = {}


def _subtimes(times, mdt, backward=False):
    """Micro-step lengths for every step of ``times`` (already directed).  Forward steps use
    ``concatenate(arange(tt, tt + dt, mdt), [tt + dt])`` (lyapunov.py:598), backward ones
    ``concatenate(arange(tt + dt, tt, mdt), [tt])`` walked in reverse (lyapunov.py:514 with
    time_direction -1), reproducing numpy's arange rounding -- which is why this is a Python loop over the steps
    (tens of milliseconds for thousands of steps); chunked runs repeat the same time vectors, so the last few
    results are kept."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    key = (times.tobytes(), float(mdt), bool(backward))
    hit = _SUBTIMES_CACHE.get(key)
    if hit is not None:
        return hit
    if len(_SUBTIMES_CACHE) >= 8:
        _SUBTIMES_CACHE.clear()
    out = _SUBTIMES_CACHE[key] = _subtimes_uncached(times, mdt, backward)
    return out


def _subtimes_uncached(times, mdt, backward):
    ptr = [0]
    subs = []
    for tt, dt in zip(times[:-1], np.diff(times)):
        if backward:
            sub = np.concatenate((np.arange(tt + dt, tt, mdt), np.full((1,), tt)))[::-1]
        else:
            sub = np.concatenate((np.arange(tt, tt + dt, mdt), np.full((1,), tt + dt)))
        d = np.diff(sub)
        subs.append(d)
        ptr.append(ptr[-1] + len(d))
    sub_dt = np.concatenate(subs) if subs else np.zeros(0)
    return np.asarray(ptr, dtype=np.int64), np.ascontiguousarray(sub_dt, dtype=np.float64)


def benettin(f, fjac, ic, mode, n_vec, q0, r0, pre_times, rec_times, mdt, write_steps, adjoint, inverse, b, c, a,
             want_r=False, want_vectors=True, seed=None, member_offset=0):
    """Run ``qgsb_lyap_benettin``.  ``pre_times`` / ``rec_times`` are the directed time vectors of the
    convergence phase and of the recorded phase.  mode 0: BLV, 1: FLV, 2: BLV whose trajectory follows the
    micro-steps (Ginelli forward pass).  ``q0 = None``: the start bases ``qr(random((n_dim, n_vec)))`` of
    lyapunov.py:592-593 are drawn on the device (``seed``: a fresh one from ``numpy.random`` when ``None``;
    ``member_offset``: global index of ``ic[0]`` when ``ic`` is a block of a larger ensemble).
    Returns ``traj (N,n,R), exp (N,m,R), vec (N,n,m,R)[, r_all]``."""
    tensor = tensor_of(f)
    if tensor_of(fjac, "fjac") is not tensor:
        raise ValueError("f and fjac must come from the same create_tendencies() call")
    ic = _lib.f64(ic)
    N, n = ic.shape
    m = int(n_vec)
    tensor.ensure_tangent(float(N) * (len(pre_times) + len(rec_times) - 2) * m)
    backward_walk = mode == 1
    ptr_a, sub_a = _subtimes(pre_times, mdt, backward_walk)
    ptr_b, sub_b = _subtimes(rec_times, mdt, backward_walk)
    sub_ptr = np.ascontiguousarray(np.concatenate((ptr_a, ptr_b[1:] + ptr_a[-1])), dtype=np.int64)
    sub_dt = np.ascontiguousarray(np.concatenate((sub_a, sub_b)))
    dt_macro = np.ascontiguousarray(np.concatenate((np.diff(pre_times), np.diff(rec_times))), dtype=np.float64)
    n_pre, n_rec = len(pre_times) - 1, len(rec_times) - 1
    R = n_records_of(rec_times, write_steps)
    b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
    if q0 is None:
        # start bases drawn and factorised on the device; the members of a sharded ensemble keep their global index
        _lib.set_seed(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64) if seed is None else seed, member_offset)
        r0 = None
    else:
        q0 = _lib.f64(q0)
        r0 = None if r0 is None else _lib.f64(r0)
    rec_traj = np.empty((N, n, R))
    rec_exp = np.empty((N, m, R))
    rec_vec = np.empty((N, n, m, R)) if want_vectors else None
    r_all = np.empty((N, n_pre + n_rec, m, m)) if want_r else None
    _lib.check(_lib.load().qgsb_lyap_benettin(
        tensor.handle, N, _lib.dptr(ic), int(mode), m, _lib.dptr(q0), _lib.dptr(r0), n_pre, n_rec,
        _lib.dptr(dt_macro), sub_ptr.ctypes.data_as(_lib.c_long_p), _lib.dptr(sub_dt), len(b), _lib.dptr(a),
        _lib.dptr(b), _lib.dptr(c), int(write_steps), 1 if adjoint else 0, float(inverse), R, _lib.dptr(rec_traj),
        _lib.dptr(rec_exp), _lib.dptr(rec_vec), _lib.dptr(r_all), None, None))
    if want_r:
        return rec_traj, rec_exp, rec_vec, r_all
    return rec_traj, rec_exp, rec_vec


def ginelli(f, fjac, ic, n_vec, q0, r0, am0, noise, noise_pert, pretime, time, aftertime, mdt, write_steps, b, c, a):
    """Run ``qgsb_clv_ginelli``: the whole of ``_compute_clv_gin_jit`` (lyapunov.py:1174-1288) on the device --
    forward Benettin pass over ``pretime``, ``time`` and ``aftertime`` keeping every ``Q`` and ``R`` in HBM, then
    the backward recursion ``A <- normalise(R^-1 A)`` from ``tc`` to ``ta``.  ``am0 (N, m, m)`` are the start
    matrices of the recursion, ``noise (N, tew, m)`` the diagonal perturbations (``None`` when ``noise_pert`` is 0).
    Returns ``traj (N,n,R), exp (N,m,R), vec (N,n,m,R)`` for the records of ``time``."""
    tensor = tensor_of(f)
    if tensor_of(fjac, "fjac") is not tensor:
        raise ValueError("f and fjac must come from the same create_tendencies() call")
    if len(aftertime) < 2:
        raise ValueError("the Ginelli method needs tc > tb")
    ic = _lib.f64(ic)
    N, n = ic.shape
    m = int(n_vec)
    tensor.ensure_tangent(float(N) * (len(pretime) + len(time) + len(aftertime) - 3) * m)
    rec_times = np.concatenate((time[:-1], aftertime))
    ptr_a, sub_a = _subtimes(pretime, mdt, False)
    ptr_b, sub_b = _subtimes(rec_times, mdt, False)
    sub_ptr = np.ascontiguousarray(np.concatenate((ptr_a, ptr_b[1:] + ptr_a[-1])), dtype=np.int64)
    sub_dt = np.ascontiguousarray(np.concatenate((sub_a, sub_b)))
    dt_macro = np.ascontiguousarray(np.concatenate((np.diff(pretime), np.diff(rec_times))), dtype=np.float64)
    dte = np.ascontiguousarray(np.concatenate((np.diff(time), np.full((1,), aftertime[1] - aftertime[0]))))
    R = n_records_of(time, write_steps)
    b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
    q0, am0 = _lib.f64(q0), _lib.f64(am0)
    r0 = None if r0 is None else _lib.f64(r0)
    noise = None if noise is None else _lib.f64(noise)
    rec_traj, rec_exp, rec_vec = np.empty((N, n, R)), np.empty((N, m, R)), np.empty((N, n, m, R))
    _lib.check(_lib.load().qgsb_clv_ginelli(
        tensor.handle, N, _lib.dptr(ic), m, _lib.dptr(q0), _lib.dptr(r0), len(pretime) - 1, len(time) - 1,
        len(aftertime) - 1, _lib.dptr(dt_macro), sub_ptr.ctypes.data_as(_lib.c_long_p), _lib.dptr(sub_dt), len(b),
        _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), int(write_steps), float(noise_pert), _lib.dptr(am0),
        _lib.dptr(noise), _lib.dptr(dte), R, _lib.dptr(rec_traj), _lib.dptr(rec_exp), _lib.dptr(rec_vec), None))
    return rec_traj, rec_exp, rec_vec


def _random_basis(n_traj, n_dim, n_vec, normal=False):
    """qr(random((n_dim, n_vec))) per member (lyapunov.py:592-593; randn for Ginelli, :1200) -- drawn and
    factorised for the whole ensemble at once (stacked LAPACK calls instead of a Python loop)."""
    draw = np.random.randn(n_traj, n_dim, n_vec) if normal else np.random.random((n_traj, n_dim, n_vec))
    q0, r0 = np.linalg.qr(draw)
    return np.ascontiguousarray(q0), np.ascontiguousarray(r0)


class _EstimatorBase(object):

    def __init__(self, num_threads=None, b=None, c=None, a=None, number_of_dimensions=None):
        if num_threads is None:
            self.num_threads = multiprocessing.cpu_count()
        else:
            self.num_threads = num_threads

        # Default is RK4
        if a is None and b is None and c is None:
            self.b, self.c, self.a = rk4_tableau()
        else:
            self.a = a
            self.b = b
            self.c = c

        self.ic = None
        self._time = None
        self._pretime = None
        self._recorded_traj = None
        self._recorded_exp = None
        self._recorded_vec = None
        self.n_traj = 0
        self.n_dim = number_of_dimensions
        self.n_records = 0
        self.n_vec = 0
        self.write_steps = 0
        self.func = None
        self.func_jac = None

    def terminate(self):
        """Stop the workers -- nothing to stop, the device context is shared (lyapunov.py:148-152)."""

    def start(self):
        """Bind to the CUDA device (the reference restarts its worker pool here, lyapunov.py:154-174)."""
        self.terminate()
        _lib.init(-1)

    def set_bca(self, b=None, c=None, a=None, ic_init=True):
        """Set the coefficients of the Runge-Kutta method and restart the estimator."""
        if a is not None:
            self.a = a
        if b is not None:
            self.b = b
        if c is not None:
            self.c = c
        if ic_init:
            self.ic = None
        self.start()

    def set_func(self, f, fjac):
        """Set the tendencies and Jacobian functions (must come from ``create_tendencies``)."""
        if tensor_of(f) is not tensor_of(fjac, "fjac"):
            raise ValueError("f and fjac must come from the same create_tendencies() call")
        self.func = f
        self.func_jac = fjac
        self.start()

    def _set_ic(self, ic):
        if ic is None:
            self.ic = np.zeros(self.func.ndim)
        else:
            self.ic = ic
        if len(self.ic.shape) == 1:
            self.ic = self.ic.reshape((1, -1))
        self.n_traj = self.ic.shape[0]
        self.n_dim = self.ic.shape[1]

    def _result(self, tt, vec):
        if self.write_steps > 0:
            if tt[::self.write_steps][-1] == tt[-1]:
                time = tt[::self.write_steps]
            else:
                time = np.concatenate((tt[::self.write_steps], np.full((1,), tt[-1])))
        else:
            time = tt[-1]
        return (time, np.squeeze(self._recorded_traj), np.squeeze(self._recorded_exp),
                None if vec is None else np.squeeze(vec))


class LyapunovsEstimator(_EstimatorBase):
    """Forward and Backward Lyapunov vectors and local exponents with the Benettin algorithm
    (reference: lyapunov.py:41-393)."""

    def __init__(self, num_threads=None, b=None, c=None, a=None, number_of_dimensions=None):
        _EstimatorBase.__init__(self, num_threads, b, c, a, number_of_dimensions)
        self._adjoint = False
        self._forward = -1
        self._inverse = 1.

    def compute_lyapunovs(self, t0, tw, t, dt, mdt, ic=None, write_steps=1, n_vec=None, forward=False, adjoint=False,
                          inverse=False, vectors=True, member_offset=0, start_basis=None):
        """Estimate the BLVs between ``tw`` and ``t`` (``forward=False``) or the FLVs between ``t0`` and ``tw``
        (``forward=True``) -- lyapunov.py:232-358.  Results via :meth:`get_lyapunovs`.  ``vectors=False`` (an
        extension) keeps only the trajectory and the local exponents: no ``(n_traj, n_dim, n_vec, n_records)`` array
        is recorded or copied, and ``get_lyapunovs`` returns ``None`` for the vectors.  ``member_offset`` (an
        extension): global index of ``ic[0]`` when this process holds one block of a sharded ensemble.
        ``start_basis`` (an extension): ``(q0 (n_traj, n_dim, n_vec), r0 (n_traj, n_vec, n_vec))`` host arrays to
        start from instead of the device-side random draw (parity tests feed the reference's own draws)."""
        if self.func is None or self.func_jac is None:
            print('No function to integrate defined!')
            return 0

        self._set_ic(ic)

        if n_vec is not None:
            self.n_vec = n_vec
        else:
            self.n_vec = self.n_dim

        self._pretime = np.concatenate((np.arange(t0, tw, dt), np.full((1,), tw)))
        self._time = np.concatenate((np.arange(tw, t, dt), np.full((1,), t)))

        self.write_steps = write_steps

        self._forward = 1 if forward else -1
        self._adjoint = adjoint
        self._inverse = 1.
        if inverse:
            self._inverse *= -1.

        # start bases: qr(random((n_dim, n_vec))) per member (lyapunov.py:592-593), drawn on the device
        q0, r0 = (None, None) if start_basis is None else start_basis
        if not forward:
            self.n_records = n_records_of(self._time, write_steps)
            res = benettin(self.func, self.func_jac, self.ic, 0, self.n_vec, q0, r0, self._pretime, self._time,
                           mdt, write_steps, adjoint, self._inverse, self.b, self.c, self.a, want_vectors=vectors,
                           member_offset=member_offset)
        else:
            self.n_records = n_records_of(self._pretime, write_steps)
            # walk back over posttime (= self._time) first, then over time (= self._pretime): lyapunov.py:509-546
            res = benettin(self.func, self.func_jac, self.ic, 1, self.n_vec, q0, r0, self._time[::-1].copy(),
                           self._pretime[::-1].copy(), mdt, write_steps, adjoint, self._inverse, self.b, self.c,
                           self.a, want_vectors=vectors, member_offset=member_offset)
        self._recorded_traj, self._recorded_exp, self._recorded_vec = res

    def get_lyapunovs(self):
        """``time, traj, exponents, vectors`` of the last estimation (lyapunov.py:360-393)."""
        tt = self._time if self._forward == -1 else self._pretime
        return self._result(tt, self._recorded_vec)


class LyapProcess(object):
    """Placeholder for the reference's worker class (lyapunov.py:396-468); replaced by CUDA thread blocks."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("LyapProcess workers were replaced by CUDA kernels; use LyapunovsEstimator")


class CovariantLyapunovsEstimator(_EstimatorBase):
    """Covariant Lyapunov vectors (reference: lyapunov.py:635-1092).

    ``method`` 0: Ginelli et al. (forward Benettin pass storing every ``Q`` and ``R`` in HBM and the backward
    triangular recursion, both on the GPU: ``qgsb_clv_ginelli``); ``method`` 1: intersection of the BLV and FLV subspaces (both
    Benettin passes and the per-record intersections on the GPU: ``qgsb_clv_subspace_intersect``).
    """

    def __init__(self, num_threads=None, b=None, c=None, a=None, number_of_dimensions=None, noise_pert=0.,
                 method=0):
        _EstimatorBase.__init__(self, num_threads, b, c, a, number_of_dimensions)
        self.noise_pert = noise_pert
        self._aftertime = None
        self._recorded_bvec = None
        self._recorded_fvec = None
        self.method = method

    def set_noise_pert(self, noise_pert):
        """Set the noise perturbation of the CLVs' diagonal during the Ginelli steps (lyapunov.py:813-826)."""
        self.noise_pert = noise_pert
        self.start()

    def compute_clvs(self, t0, ta, tb, tc, dt, mdt, ic=None, write_steps=1, n_vec=None, method=None,
                     backward_vectors=False, forward_vectors=False):
        """Estimate the CLVs between ``ta`` and ``tb`` (lyapunov.py:863-988).  Results via :meth:`get_clvs`."""
        if self.func is None or self.func_jac is None:
            print('No function to integrate defined!')
            return 0

        self._set_ic(ic)

        if n_vec is not None:
            self.n_vec = n_vec
        else:
            self.n_vec = self.n_dim

        if method is not None:
            self.method = method

        self._pretime = np.concatenate((np.arange(t0, ta, dt), np.full((1,), ta)))
        self._time = np.concatenate((np.arange(ta, tb, dt), np.full((1,), tb)))
        self._aftertime = np.concatenate((np.arange(tb, tc, dt), np.full((1,), tc)))

        self.write_steps = write_steps
        self.n_records = n_records_of(self._time, write_steps)
        self._recorded_bvec = None
        self._recorded_fvec = None

        if self.method == 0:
            self._ginelli(mdt)
        else:
            self._subspaces(mdt, backward_vectors, forward_vectors)

    # ---- method 0: Ginelli et al., lyapunov.py:1174-1288 ---------------------------------------------
    def _ginelli(self, mdt):
        n_traj, n_dim, n_vec = self.n_traj, self.n_dim, self.n_vec
        tew = len(self._time) + len(self._aftertime) - 2
        q0, r0 = _random_basis(n_traj, n_dim, n_vec, normal=True)
        # start of the backward recursion (lyapunov.py:1254-1255) and the diagonal noise of every step (:1260, :1271)
        am0 = np.empty((n_traj, n_vec, n_vec))
        for i in range(n_traj):
            am0[i], _ = normalize_matrix_columns(np.linalg.qr(np.random.randn(n_dim, n_vec))[1])
        noise = np.random.randn(n_traj, tew, n_vec) if self.noise_pert != 0. else None
        res = ginelli(self.func, self.func_jac, self.ic, n_vec, q0, r0, am0, noise, self.noise_pert, self._pretime,
                      self._time, self._aftertime, mdt, self.write_steps, self.b, self.c, self.a)
        self._recorded_traj, self._recorded_exp, self._recorded_vec = res

    # ---- method 1: subspace intersection, lyapunov.py:1292-1329 ----------------------------------------
    def _subspaces(self, mdt, backward_vectors, forward_vectors):
        n_traj, n_dim = self.n_traj, self.n_dim
        self.n_vec = n_dim
        pretime, time, aftertime = self._pretime, self._time, self._aftertime
        # state at ta, start of the FLV walk (the reference stores the whole trajectory, :1299)
        ic_a = _integrate_runge_kutta_jit(self.func, pretime, self.ic, 1, 0, self.b, self.c, self.a)[:, :, -1]
        q0, r0 = _random_basis(n_traj, n_dim, n_dim)
        traj, exp, fvec = benettin(self.func, self.func_jac, ic_a, 1, n_dim, q0, r0, aftertime[::-1].copy(),
                                   time[::-1].copy(), mdt, self.write_steps, False, 1., self.b, self.c, self.a)
        q0, r0 = _random_basis(n_traj, n_dim, n_dim)
        traj, exp, bvec = benettin(self.func, self.func_jac, self.ic, 0, n_dim, q0, r0, pretime, time, mdt,
                                   self.write_steps, False, 1., self.b, self.c, self.a)
        n_records = traj.shape[-1]
        # intersection of the BLV / FLV subspaces of every record (lyapunov.py:1315-1320) on the device
        recorded_vec = np.empty((n_traj, n_dim, n_dim, n_records))
        _lib.check(_lib.load().qgsb_clv_subspace_intersect(n_traj, n_dim, n_records, _lib.dptr(_lib.f64(bvec)),
                                                           _lib.dptr(_lib.f64(fvec)), _lib.dptr(recorded_vec)))
        # local exponents: one micro-step of the tangent model on every (member, record) pair (:1322-1327)
        subtime = np.array([0., mdt])
        ys = np.ascontiguousarray(np.moveaxis(traj, 2, 1).reshape(n_traj * n_records, n_dim))
        vecs = np.ascontiguousarray(np.moveaxis(recorded_vec, 3, 1).reshape(n_traj * n_records, n_dim, n_dim))
        _, sol = _integrate_runge_kutta_tgls_jit(self.func, self.func_jac, subtime, ys, vecs, 1, 0, self.b, self.c,
                                                 self.a, False, 1., _zeros_func)
        norms = np.linalg.norm(sol[:, :, :, 0], 2, axis=1).reshape(n_traj, n_records, n_dim)
        self._recorded_traj = traj
        self._recorded_exp = np.moveaxis(np.log(np.abs(norms)) / mdt, 1, 2)
        self._recorded_vec = recorded_vec
        if forward_vectors:
            self._recorded_fvec = fvec
        if backward_vectors:
            self._recorded_bvec = bvec

    def get_clvs(self):
        """``time, traj, exponents, vectors`` of the last estimation (lyapunov.py:990-1018)."""
        return self._result(self._time, self._recorded_vec)

    def get_blvs(self):
        """BLVs obtained during the last estimation -- only with ``method=1`` (lyapunov.py:1020-1055)."""
        if self._recorded_bvec is None:
            return None
        return self._result(self._time, self._recorded_bvec)

    def get_flvs(self):
        """FLVs obtained during the last estimation -- only with ``method=1`` (lyapunov.py:1057-1092)."""
        if self._recorded_fvec is None:
            return None
        return self._result(self._time, self._recorded_fvec)


class ClvProcess(object):
    """Placeholder for the reference's worker class (lyapunov.py:1095-1170); replaced by CUDA thread blocks."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("ClvProcess workers were replaced by CUDA kernels; use CovariantLyapunovsEstimator")
