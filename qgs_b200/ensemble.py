"""Device-resident ensembles, member sharding across ranks and ensemble statistics.

Two callers sit either side of the integrate kernels in the reference and are the "next" rows of
SURVEY.md section 8f:

* the chunked-run idiom of ``qgs_maooam.py:115-136`` / ``RungeKuttaIntegrator.initialize``
  (integrator.py:198-295) feeds the final state of one ``integrate(..., write_steps=0)`` call back as
  the initial condition of the next; :class:`DeviceEnsemble` keeps that state in HBM between calls;
* ``TrajectoriesStatistics.compute_stats`` (qgs/integrators/statistics.py:33-66) averages functions
  of the trajectories over the ensemble; :meth:`DeviceEnsemble.moments` reduces the first two moments
  on the device and, when the members are sharded over ranks (one process per GPU), sums the
  ``2 * n_dim`` partial sums with one ``torch.distributed`` all-reduce -- the only collective of the
  path (NCCL over NVLink on GPUs, gloo in the CPU tests).

Members are independent ODE solves (integrator.py:388-389): rank ``g`` of ``G`` owns the contiguous
block ``shard_bounds(N, G, g)`` and nothing is exchanged during integration.
"""
import ctypes

import numpy as np

from qgs_b200 import _lib
from qgs_b200.integrators.integrate import directed_dt, rk4_tableau, tensor_of


def shard_bounds(n_traj, world_size, rank):
    """Balanced contiguous partition of the members: rank ``g`` gets ``[g * N // G, (g + 1) * N // G)`` -- block sizes
    differ by at most one and no rank is left empty when ``N >= G`` (the library's own split across the devices of
    one process, ``qgsb::shard_range``, is the same rule)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    n_traj = int(n_traj)
    return rank * n_traj // world_size, (rank + 1) * n_traj // world_size


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def all_reduce_sums(local, device=None):
    """Sum a small float64 vector over all ranks (no-op without an initialised process group)."""
    dist = _dist()
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    t = torch.from_numpy(local.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda() if device is None else t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def combine_moments(local_sum, local_sumsq, local_count):
    """Global mean and (population) variance per variable from per-rank partial sums."""
    n = len(local_sum)
    packed = np.concatenate((local_sum, local_sumsq, [float(local_count)]))
    tot = all_reduce_sums(packed)
    count = tot[-1]
    mean = tot[:n] / count
    var = np.maximum(tot[n:2 * n] / count - mean * mean, 0.)
    return mean, var, int(round(count))


def gather_states(local_states):
    """All ranks' final states concatenated in rank order (the optional final gather of SURVEY.md section 8e)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local_states
    import torch
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.ascontiguousarray(local_states))
    return np.concatenate(parts, axis=0)


class DeviceEnsemble(object):
    """Ensemble state resident in HBM (tiled structure-of-arrays, see ``include/qgsb.h``).

    ``f`` is a tendencies callable from ``create_tendencies``.  ``ic`` is ``(n_traj, n_dim)``; with
    ``sharded=True`` and an initialised ``torch.distributed`` group only this rank's block of members
    is uploaded.
    """

    def __init__(self, f, ic, sharded=False):
        self.tensor = tensor_of(f)
        ic = np.atleast_2d(np.asarray(ic, dtype=np.float64))
        if ic.shape[1] != self.tensor.ndim:
            raise ValueError("ic must have shape (n_traj, %d)" % self.tensor.ndim)
        self.n_global = ic.shape[0]
        self.lo, self.hi = 0, ic.shape[0]
        dist = _dist()
        if sharded and dist is not None:
            # the same test on every rank, before anything collective: no rank may go on alone into an all-reduce
            if ic.shape[0] < dist.get_world_size():
                raise ValueError("%d members cannot be sharded over %d ranks" % (ic.shape[0], dist.get_world_size()))
            self.lo, self.hi = shard_bounds(ic.shape[0], dist.get_world_size(), dist.get_rank())
        local = _lib.f64(ic[self.lo:self.hi])
        if local.shape[0] == 0:
            raise ValueError("an ensemble needs at least one member")
        self.n_traj = local.shape[0]
        self.n_dim = self.tensor.ndim
        self.time = 0.
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().qgsb_ensemble_create(self.tensor.handle, self.n_traj, ctypes.byref(handle)))
        self._handle = handle
        _lib.check(_lib.load().qgsb_ensemble_upload(self._handle, _lib.dptr(local)))

    def close(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().qgsb_ensemble_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self.close()

    def integrate(self, t0, t, dt, forward=True, b=None, c=None, a=None):
        """Advance every member from ``t0`` to ``t`` (``write_steps=0`` semantics) without leaving the device.
        Returns the kernel time in milliseconds."""
        if a is None and b is None and c is None:
            b, c, a = rk4_tableau()
        time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))
        steps = directed_dt(time, 1 if forward else -1)
        b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
        ms = ctypes.c_double()
        _lib.check(_lib.load().qgsb_ensemble_integrate(self._handle, len(steps), _lib.dptr(steps), len(b),
                                                       _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ctypes.byref(ms)))
        self.time = time[-1] if forward else time[0]
        return ms.value

    def integrate_trajectories(self, t0, t, dt, forward=True, write_steps=1, b=None, c=None, a=None, out=None):
        """Advance every member from ``t0`` to ``t`` and return ``(time, traj)`` exactly as
        ``RungeKuttaIntegrator.integrate`` + ``get_trajectories`` would (``traj`` ``(n_local, n_dim, n_records)``, not
        squeezed), but from the state resident on the device -- no initial conditions are uploaded -- and with the
        records streamed to the host in chunks while the integration goes on, so device memory does not grow with
        the number of records.  ``out`` may be a preallocated (e.g. pinned) array."""
        if a is None and b is None and c is None:
            b, c, a = rk4_tableau()
        time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))
        steps = directed_dt(time, 1 if forward else -1)
        ws = int(write_steps)
        rec_time = self._record_times(time, ws, forward)
        R = len(rec_time)
        b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
        if out is None:
            out = np.empty((self.n_traj, self.n_dim, R))
        if out.shape != (self.n_traj, self.n_dim, R) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % ((self.n_traj, self.n_dim, R),))
        ms = ctypes.c_double()
        _lib.check(_lib.load().qgsb_ensemble_integrate_trajectories(
            self._handle, len(steps), _lib.dptr(steps), len(b), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ws,
            1 if forward else -1, R, _lib.dptr(out), ctypes.byref(ms)))
        self.time = time[-1] if forward else time[0]
        self.last_ms = ms.value
        return rec_time, out

    @staticmethod
    def _record_times(time, ws, forward):
        """Returned time vector of integrator.py:409-424 (for write_steps == 0 the reference returns time[-1] in both
        directions)."""
        if ws > 0 and forward:
            rec_time = time[::ws]
            return rec_time if rec_time[-1] == time[-1] else np.concatenate((rec_time, time[-1:]))
        if ws > 0:
            rec_time = time[::-ws][::-1]
            return rec_time if rec_time[0] == time[0] else np.concatenate((time[:1], rec_time))
        return time[-1:]

    def integrate_moments(self, t0, t, dt, forward=True, write_steps=1, b=None, c=None, a=None):
        """Advance every member from ``t0`` to ``t`` and return ``(time, mean, var)`` of the ensemble at every
        record the reference would write (integrate.py:190-221, integrator.py:397-424): ``mean`` and ``var`` have
        shape ``(n_records, n_dim)`` and cover ALL ranks' members.  This is what
        ``TrajectoriesStatistics.compute_stats`` (qgs/integrators/statistics.py:33-66) yields for the functions
        ``x`` and ``x**2``, but the ``(n_traj, n_dim, n_records)`` trajectories never leave the device: records
        are reduced in HBM and only ``2 * n_records * n_dim`` sums reach the host (and the all-reduce)."""
        if a is None and b is None and c is None:
            b, c, a = rk4_tableau()
        time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))
        steps = directed_dt(time, 1 if forward else -1)
        n_steps = len(steps)
        ws = int(write_steps)
        rec_time = self._record_times(time, ws, forward)
        R = len(rec_time)
        b, c, a = _lib.f64(b), _lib.f64(c), _lib.f64(a)
        s1, s2 = np.empty((R, self.n_dim)), np.empty((R, self.n_dim))
        ms = ctypes.c_double()
        _lib.check(_lib.load().qgsb_ensemble_integrate_moments(self._handle, n_steps, _lib.dptr(steps), len(b),
                                                               _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ws, R,
                                                               _lib.dptr(s1), _lib.dptr(s2), ctypes.byref(ms)))
        self.time = time[-1] if forward else time[0]
        self.last_ms = ms.value
        packed = np.concatenate((s1.ravel(), s2.ravel(), [float(self.n_traj)]))
        tot = all_reduce_sums(packed)
        count = tot[-1]
        mean = tot[:R * self.n_dim].reshape(R, self.n_dim) / count
        var = np.maximum(tot[R * self.n_dim:-1].reshape(R, self.n_dim) / count - mean * mean, 0.)
        if not forward and ws > 0:
            # the kernel records in integration order; the reference returns the record axis in increasing time
            mean, var = mean[::-1], var[::-1]
        return rec_time, mean, var

    def set_states(self, states):
        """Replace the resident members by ``states`` ``(n_local, n_dim)`` (same number of members)."""
        states = _lib.f64(np.atleast_2d(states))
        if states.shape != (self.n_traj, self.n_dim):
            raise ValueError("states must have shape %s" % ((self.n_traj, self.n_dim),))
        _lib.check(_lib.load().qgsb_ensemble_upload(self._handle, _lib.dptr(states)))

    def states(self):
        """This rank's members as a host array ``(n_local, n_dim)``."""
        out = np.empty((self.n_traj, self.n_dim))
        _lib.check(_lib.load().qgsb_ensemble_download(self._handle, _lib.dptr(out)))
        return out

    def local_sums(self):
        s1, s2 = np.empty(self.n_dim), np.empty(self.n_dim)
        _lib.check(_lib.load().qgsb_ensemble_moments(self._handle, _lib.dptr(s1), _lib.dptr(s2)))
        return s1, s2

    def moments(self):
        """Ensemble mean and variance of every variable over ALL ranks' members."""
        s1, s2 = self.local_sums()
        mean, var, _ = combine_moments(s1, s2, self.n_traj)
        return mean, var


def spectrum_from_local_exponents(local_exp, n_global=None):
    """Ensemble- and time-mean Lyapunov spectrum with its standard error from the local exponents
    ``(n_local, n_vec, n_records)`` of THIS rank's members (``LyapunovsEstimator.get_lyapunovs()[2]``); the
    partial sums of all ranks are combined with one all-reduce of ``2 * n_vec + 1`` doubles.  The standard error
    is over members (each member's time mean is one sample)."""
    local_exp = np.asarray(local_exp, dtype=np.float64)
    if local_exp.ndim == 2:
        local_exp = local_exp[None]
    per_member = local_exp.mean(axis=2)                       # (n_local, n_vec)
    packed = np.concatenate((per_member.sum(axis=0), (per_member ** 2).sum(axis=0), [float(per_member.shape[0])]))
    tot = all_reduce_sums(packed)
    m = local_exp.shape[1]
    count = tot[-1]
    mean = tot[:m] / count
    var = np.maximum(tot[m:2 * m] / count - mean * mean, 0.)
    sem = np.sqrt(var / max(count - 1., 1.))
    if n_global is not None and int(round(count)) != int(n_global):
        raise RuntimeError("spectrum combined %d members, expected %d" % (int(round(count)), int(n_global)))
    return mean, sem


def sharded_lyapunov_spectrum(f, Df, ic, t0, tw, t, dt, mdt, n_vec=None, write_steps=1, forward=False):
    """Lyapunov spectrum of an ensemble sharded over the ranks of the current ``torch.distributed`` group
    (BASELINE.json config "MAOOAM-36 Lyapunov spectrum ..., 8-GPU sharded ensemble").  Every rank runs
    ``LyapunovsEstimator.compute_lyapunovs`` (lyapunov.py:232-358) on ITS block of ``ic`` -- members are
    independent and carry their own basis, so nothing is exchanged during the Benettin loop -- and the
    member/time-mean exponents are combined at the end.  Returns ``(mean, standard_error, local_result)`` where
    ``local_result`` is this rank's ``get_lyapunovs()`` tuple."""
    from qgs_b200.toolbox.lyapunov import LyapunovsEstimator
    ic = np.atleast_2d(np.asarray(ic, dtype=np.float64))
    dist = _dist()
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist is not None else (1, 0)
    if ic.shape[0] < world:      # the same test on every rank: nobody enters the all-reduce alone
        raise ValueError("%d members cannot be sharded over %d ranks" % (ic.shape[0], world))
    lo, hi = shard_bounds(ic.shape[0], world, rank)
    est = LyapunovsEstimator()
    est.set_func(f, Df)
    est.compute_lyapunovs(t0, tw, t, dt, mdt, ic=ic[lo:hi], write_steps=write_steps, n_vec=n_vec, forward=forward,
                          vectors=False, member_offset=lo)
    res = est.get_lyapunovs()
    exps = np.asarray(res[2])
    m = est.n_vec
    exps = exps.reshape(hi - lo, m, -1)
    mean, sem = spectrum_from_local_exponents(exps, ic.shape[0])
    return mean, sem, res
