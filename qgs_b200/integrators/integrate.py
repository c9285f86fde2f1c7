"""Functional integrators on the GPU -- mirror of ``qgs/integrators/integrate.py``.

``integrate_runge_kutta`` and ``integrate_runge_kutta_tgls`` keep the reference's signatures,
defaults, shape conventions and returned time vectors (integrate.py:29-179, :240-552).  The jitted
loops ``_integrate_runge_kutta_jit`` / ``_integrate_runge_kutta_tgls_jit`` become one fused kernel
launch over all members, stages and steps (``qgsb_rk_integrate`` / ``qgsb_rk_tgls_integrate``).

``f`` (and ``fjac``) must be the callables returned by
:func:`qgs_b200.functions.tendencies.create_tendencies` or ``tendencies_from_tensor``: they carry the
device-resident tensor.  Arbitrary Python/numba callables cannot run inside a CUDA kernel and are
rejected -- there is no CPU fallback.
"""
import ctypes

import numpy as np

from qgs_b200 import _lib
from qgs_b200.functions.tendencies import Tendencies, JacobianTendencies
from qgs_b200.functions.util import reverse


def _zeros_func(t, x):
    """The reference's default inhomogeneous term (integrate.py:235-237); the only one supported on the device."""
    return np.zeros_like(x)


def rk4_tableau():
    """Classic RK4 coefficients, the reference default (integrate.py:149-155)."""
    c = np.array([0., 0.5, 0.5, 1.])
    b = np.array([1. / 6, 1. / 3, 1. / 3, 1. / 6])
    a = np.zeros((len(c), len(b)))
    a[1, 0] = 0.5
    a[2, 1] = 0.5
    a[3, 2] = 1.
    return b, c, a


def tensor_of(f, what="f"):
    if isinstance(f, (Tendencies, JacobianTendencies)):
        return f.tensor
    raise TypeError("%s must be a callable returned by qgs_b200 create_tendencies()/tendencies_from_tensor() "
                    "(it carries the device tensor); got %r.  Arbitrary Python or numba functions cannot be "
                    "integrated by the CUDA kernels and there is no CPU fallback." % (what, type(f)))


def check_boundary(boundary):
    if boundary is not None and boundary is not _zeros_func:
        raise NotImplementedError("only the homogeneous tangent linear model (boundary=None) runs on the device")


def n_records_of(time, write_steps):
    """integrate.py:190-196."""
    if write_steps == 0:
        return 1
    tot = time[::write_steps]
    n_records = len(tot)
    if tot[-1] != time[-1]:
        n_records += 1
    return n_records


def directed_dt(time, time_direction):
    """np.diff(directed_time) of integrate.py:199-205: the per-step (signed) step lengths."""
    directed_time = reverse(time) if time_direction == -1 else np.asarray(time, dtype=np.float64)
    return np.ascontiguousarray(np.diff(directed_time))


def _integrate_runge_kutta_jit(f, time, ic, time_direction, write_steps, b, c, a):
    """Device replacement of integrate.py:182-223.  Returns ``(n_traj, n_dim, n_records)``."""
    tensor = tensor_of(f)
    time = _lib.f64(time)
    ic = _lib.f64(ic)
    if ic.ndim != 2 or ic.shape[1] != tensor.ndim:
        raise ValueError("ic must have shape (n_traj, %d), got %s" % (tensor.ndim, ic.shape))
    b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
    dt = directed_dt(time, time_direction)
    n_rec = n_records_of(time, write_steps)
    traj = np.empty((ic.shape[0], ic.shape[1], n_rec))
    _lib.check(_lib.load().qgsb_rk_integrate(tensor.handle, ic.shape[0], _lib.dptr(ic), len(dt), _lib.dptr(dt),
                                             len(b), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), int(write_steps),
                                             int(time_direction), n_rec, _lib.dptr(traj), None))
    return traj


def _integrate_runge_kutta_tgls_jit(f, fjac, time, ic, tg_ic, time_direction, write_steps, b, c, a,
                                    adjoint, inverse, boundary):
    """Device replacement of integrate.py:555-614.  Returns ``(n_traj, n_dim, R)``, ``(n_traj, n_dim, m, R)``."""
    tensor = tensor_of(f)
    if tensor_of(fjac, "fjac") is not tensor:
        raise ValueError("f and fjac must come from the same create_tendencies() call")
    check_boundary(boundary)
    time = _lib.f64(time)
    ic = _lib.f64(ic)
    tg_ic = _lib.f64(tg_ic)
    if ic.ndim != 2 or ic.shape[1] != tensor.ndim:
        raise ValueError("ic must have shape (n_traj, %d), got %s" % (tensor.ndim, ic.shape))
    if tg_ic.ndim != 3 or tg_ic.shape[0] != ic.shape[0] or tg_ic.shape[1] != ic.shape[1]:
        raise ValueError("tg_ic must have shape (n_traj, n_dim, n_tg_traj), got %s" % (tg_ic.shape,))
    tensor.ensure_tangent(float(ic.shape[0]) * max(len(time) - 1, 0) * tg_ic.shape[2])
    b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
    dt = directed_dt(time, time_direction)
    n_rec = n_records_of(time, write_steps)
    m = tg_ic.shape[2]
    traj = np.empty((ic.shape[0], ic.shape[1], n_rec))
    fmat = np.empty((ic.shape[0], ic.shape[1], m, n_rec))
    _lib.check(_lib.load().qgsb_rk_tgls_integrate(
        tensor.handle, ic.shape[0], _lib.dptr(ic), m, _lib.dptr(tg_ic), len(dt), _lib.dptr(dt), len(b), _lib.dptr(a),
        _lib.dptr(b), _lib.dptr(c), int(write_steps), int(time_direction), 1 if adjoint else 0, float(inverse), n_rec,
        _lib.dptr(traj), _lib.dptr(fmat), None))
    return traj, fmat


def infer_ic(f):
    """Zero initial condition of the right size (integrate.py:131-143)."""
    tensor_of(f)
    return np.zeros(f.ndim)


def normalise_tg_ic(tg_ic, n_traj, n_dim):
    """tg_ic shape rules of integrate.py:479-500 / integrator.py:933-958 -> (n_traj, n_dim, n_tg_traj)."""
    if len(tg_ic.shape) == 1:
        tg_ic = tg_ic.reshape((1, -1, 1))
        ict = tg_ic.copy()
        for i in range(n_traj - 1):
            ict = np.concatenate((ict, tg_ic))
        tg_ic = ict
    elif len(tg_ic.shape) == 2:
        if tg_ic.shape[0] == n_traj:
            tg_ic = tg_ic[..., np.newaxis]
        else:
            tg_ic = tg_ic[np.newaxis, ...]
            tg_ic = np.swapaxes(tg_ic, 1, 2)
            ict = tg_ic.copy()
            for i in range(n_traj - 1):
                ict = np.concatenate((ict, tg_ic))
            tg_ic = ict
    elif len(tg_ic.shape) == 3:
        if tg_ic.shape[1] != n_dim:
            tg_ic = np.swapaxes(tg_ic, 1, 2)
    return tg_ic


def restore_fmatrix_orientation(recorded_fmatrix, tg_ic_sav, n_dim):
    """Swap the fundamental matrix back to the caller's orientation (integrate.py:527-534)."""
    if len(tg_ic_sav.shape) == 2:
        if recorded_fmatrix.shape[1:3] != tg_ic_sav.shape:
            recorded_fmatrix = np.swapaxes(recorded_fmatrix, 1, 2)
    elif len(tg_ic_sav.shape) == 3:
        if tg_ic_sav.shape[1] != n_dim:
            if recorded_fmatrix.shape[:3] != tg_ic_sav.shape:
                recorded_fmatrix = np.swapaxes(recorded_fmatrix, 1, 2)
    return recorded_fmatrix


def returned_time(time, t0, t, forward, write_steps):
    """The time vector rules of integrate.py:166-179."""
    if write_steps > 0:
        if forward:
            if time[::write_steps][-1] == time[-1]:
                return time[::write_steps]
            return np.concatenate((time[::write_steps], np.full((1,), t)))
        rtime = reverse(time[::-write_steps])
        if rtime[0] == time[0]:
            return rtime
        return np.concatenate((np.full((1,), t0), rtime))
    return time[-1]


def integrate_runge_kutta(f, t0, t, dt, ic=None, forward=True, write_steps=1, b=None, c=None, a=None):
    """Integrate ``dx/dt = f(t, x)`` with an explicit Runge-Kutta method (integrate.py:29-179).

    Same arguments and returns as the reference: ``time, traj`` with ``traj`` of shape
    ``(n_traj, n_dim, n_steps)`` squeezed.
    """
    if ic is None:
        ic = infer_ic(f)

    if len(ic.shape) == 1:
        ic = ic.reshape((1, -1))

    # Default is RK4
    if a is None and b is None and c is None:
        b, c, a = rk4_tableau()

    time_direction = 1 if forward else -1

    time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))

    recorded_traj = _integrate_runge_kutta_jit(f, time, ic, time_direction, write_steps, b, c, a)

    return returned_time(time, t0, t, forward, write_steps), np.squeeze(recorded_traj)


def integrate_runge_kutta_tgls(f, fjac, t0, t, dt, ic=None, tg_ic=None,
                               forward=True, adjoint=False, inverse=False, boundary=None,
                               write_steps=1, b=None, c=None, a=None):
    """Integrate the model and its tangent linear (or adjoint) model together (integrate.py:240-552).

    Same arguments, ``tg_ic`` shape rules and returns as the reference: ``time, traj, tg_traj``.
    """
    if ic is None:
        ic = infer_ic(f)

    if len(ic.shape) == 1:
        ic = ic.reshape((1, -1))

    n_traj = ic.shape[0]

    if tg_ic is None:
        tg_ic = np.eye(ic.shape[1])

    tg_ic_sav = tg_ic.copy()
    tg_ic = normalise_tg_ic(tg_ic, n_traj, ic.shape[1])

    # Default is RK4
    if a is None and b is None and c is None:
        b, c, a = rk4_tableau()

    time_direction = 1 if forward else -1

    time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))

    inv = 1.
    if inverse:
        inv *= -1.

    recorded_traj, recorded_fmatrix = _integrate_runge_kutta_tgls_jit(f, fjac, time, ic, tg_ic,
                                                                      time_direction, write_steps,
                                                                      b, c, a, adjoint, inv, boundary)

    recorded_fmatrix = restore_fmatrix_orientation(recorded_fmatrix, tg_ic_sav, ic.shape[1])

    return returned_time(time, t0, t, forward, write_steps), np.squeeze(recorded_traj), np.squeeze(recorded_fmatrix)
