"""Ensemble integrator classes on the GPU -- mirror of ``qgs/integrators/integrator.py``.

``RungeKuttaIntegrator`` and ``RungeKuttaTglsIntegrator`` keep the reference's constructor,
methods, attributes, shape conventions and quirks (integrator.py:27-450, :515-1100).  What is gone is
the pool of ``multiprocessing`` workers fed one pickled message per trajectory
(integrator.py:121-142, 388-395): ``integrate`` hands the whole ensemble to one fused CUDA launch.
``num_threads`` is kept for signature parity; ``initialize`` batches by the size of the device, not by it.
"""
import multiprocessing

import numpy as np

from qgs_b200 import _lib
from qgs_b200.functions.util import reverse
from qgs_b200.integrators.integrate import (_integrate_runge_kutta_jit, _integrate_runge_kutta_tgls_jit, _zeros_func,
                                            check_boundary, n_records_of, normalise_tg_ic, rk4_tableau, tensor_of)


class _IntegratorBase(object):

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
        self._recorded_traj = None
        self.n_traj = 0
        self.n_dim = number_of_dimensions
        self.n_records = 0
        self._write_steps = 0
        self._time_direction = 1
        self.func = None
        self._started = False

    def terminate(self):
        """Release the resources of the integrator.  The reference kills its worker processes here
        (integrator.py:113-119); the device context is shared by all integrators of the process, so there
        is nothing to kill.  Safe to call on a never-started integrator, like the reference."""
        self._started = False

    def start(self):
        """Bind to the CUDA device (the reference (re)starts its worker pool here, integrator.py:121-142)."""
        self.terminate()
        _lib.init(-1)
        self._started = True

    def set_bca(self, b=None, c=None, a=None, ic_init=True):
        """Set the coefficients of the Runge-Kutta method (integrator.py:167-196)."""
        if a is not None:
            self.a = a
        if b is not None:
            self.b = b
        if c is not None:
            self.c = c
        if ic_init:
            self.ic = None
        self.start()

    def _infer_ndim(self):
        # integrator.py:344-361 probes f with growing zero vectors; our f knows its dimension
        if self.n_dim is not None:
            return self.n_dim
        return len(self.func(0., np.zeros(self.func.ndim)))

    def _returned_time(self):
        if self._write_steps > 0:
            if self._time_direction == 1:
                if self._time[::self._write_steps][-1] == self._time[-1]:
                    return self._time[::self._write_steps]
                return np.concatenate((self._time[::self._write_steps], np.full((1,), self._time[-1])))
            rtime = reverse(self._time[::-self._write_steps])
            if rtime[0] == self._time[0]:
                return rtime
            return np.concatenate((np.full((1,), self._time[0]), rtime))
        return self._time[-1]

    def get_ic(self):
        """Returns the initial conditions stored in the integrator."""
        return self.ic

    def set_ic(self, ic):
        """Direct setter for the integrator's initial conditions."""
        self.ic = ic

    def _prepare(self, t0, t, dt, ic, forward, write_steps):
        if ic is None:
            if self.ic is None:
                self.ic = np.zeros(self._infer_ndim())
        else:
            self.ic = ic

        if len(self.ic.shape) == 1:
            self.ic = self.ic.reshape((1, -1))

        self.n_traj = self.ic.shape[0]
        self.n_dim = self.ic.shape[1]
        self._time = np.concatenate((np.arange(t0, t, dt), np.full((1,), t)))
        self._write_steps = write_steps
        self._time_direction = 1 if forward else -1
        self.n_records = n_records_of(self._time, write_steps)


class RungeKuttaIntegrator(_IntegratorBase):
    """Integrate ``dx/dt = f(t, x)`` for an ensemble of initial conditions with an explicit Runge-Kutta method
    (reference: integrator.py:27-450).

    Parameters and attributes are the reference's: ``num_threads, b, c, a, number_of_dimensions``;
    ``n_dim, n_traj, n_records, ic, func``.
    """

    def set_func(self, f, ic_init=True):
        """Set the tendencies function to integrate; must come from ``create_tendencies`` (it carries the device
        tensor).  Restarts the integrator like the reference (integrator.py:144-165)."""
        tensor_of(f)
        self.func = f
        if ic_init:
            self.ic = None
        self.start()

    #: members per batch of :meth:`initialize` when it reconverges perturbed copies; ``None``: one full wave of the
    #: fused kernel (2 thread blocks of 128 members on every SM).  ``num_threads`` only sets a lower bound.
    device_batch = None

    def _initialize_batch(self):
        if self.device_batch is not None:
            return max(int(self.device_batch), 1)
        return max(int(self.num_threads), 2 * 128 * _lib.device_info()["sm_count"])

    def initialize(self, convergence_time, dt, pert_size=0.01, reconvergence_time=None, forward=True,
                   number_of_trajectories=1, ic=None, reconverge=False):
        """Put ``number_of_trajectories`` members on the attractor (integrator.py:198-295) and store them in ``ic``.

        Same arguments and the same scheme as the reference -- a first batch of members is integrated over
        ``convergence_time``; with ``reconverge`` (forced when more members are asked for than one batch holds) every
        further batch is the previous one plus a perturbation of size ``pert_size``, integrated over the shorter
        ``reconvergence_time`` -- but the batch is sized for the device, not for the host: the reference's batch is
        ``num_threads`` worker processes (integrator.py:257-291), here it is one full wave of thread blocks
        (:attr:`device_batch`; ``num_threads`` is only a lower bound), so a million members are a few dozen launches
        instead of tens of thousands.  The members never travel as trajectories: each batch lives in a resident
        ensemble (``qgsb_ensemble_*``), only its end state is read back to be perturbed for the next batch."""
        from qgs_b200.ensemble import DeviceEnsemble
        if self.func is None:
            print('No function to integrate defined!')
            return 0
        reconverge = bool(reconverge)
        batch = self._initialize_batch()

        if ic is None:
            n_dim = self._infer_ndim()
            n_total = int(number_of_trajectories)
            if n_total > batch:
                reconverge = True
            first = np.random.randn(min(n_total, batch) if reconverge else n_total, n_dim)
        else:
            first = np.atleast_2d(np.array(ic, dtype=np.float64))
            n_total = first.shape[0]
            n_dim = first.shape[1]
            if reconverge and reconvergence_time is not None:
                first = first[:batch]          # the reference keeps the first batch only (integrator.py:236-238)

        ens = DeviceEnsemble(self.func, first)
        try:
            ens.integrate(0., convergence_time, dt, forward=forward, b=self.b, c=self.c, a=self.a)
            x = ens.states()
            if not (reconverge and reconvergence_time is not None) or x.shape[0] >= n_total:
                out = x
            else:
                out = np.empty((n_total, n_dim))
                out[:x.shape[0]] = x
                index = x.shape[0]
                while index < n_total:
                    count = min(x.shape[0], n_total - index)
                    x = x[:count] + pert_size * np.random.randn(count, n_dim)
                    if count != ens.n_traj:
                        ens.close()
                        ens = DeviceEnsemble(self.func, x)
                    else:
                        ens.set_states(x)
                    ens.integrate(0., reconvergence_time, dt, forward=forward, b=self.b, c=self.c, a=self.a)
                    x = ens.states()
                    out[index:index + count] = x
                    index += count
        finally:
            ens.close()
        self.n_traj = out.shape[0]
        self.n_dim = n_dim
        # like the reference, ic ends up as what get_trajectories() returns: squeezed for a single member
        self.ic = np.squeeze(out) if out.shape[0] == 1 else out

    def integrate(self, t0, t, dt, ic=None, forward=True, write_steps=1):
        """Integrate the ensemble from ``t0`` to ``t`` (integrator.py:297-395).  Results via ``get_trajectories``."""
        if self.func is None:
            print('No function to integrate defined!')
            return 0

        self._prepare(t0, t, dt, ic, forward, write_steps)
        self._recorded_traj = _integrate_runge_kutta_jit(self.func, self._time, self.ic, self._time_direction,
                                                         self._write_steps, self.b, self.c, self.a)

    def get_trajectories(self):
        """``time, traj`` of the last integration (integrator.py:397-424); ``traj`` is
        ``(n_traj, n_dim, n_records)`` squeezed."""
        return self._returned_time(), np.squeeze(self._recorded_traj)


class TrajectoryProcess(object):
    """Placeholder for the reference's worker process class (integrator.py:453-512).  Workers no longer
    exist: trajectories are integrated by CUDA thread blocks."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("TrajectoryProcess workers were replaced by CUDA kernels; use RungeKuttaIntegrator")


class RungeKuttaTglsIntegrator(_IntegratorBase):
    """Integrate the model together with its tangent linear or adjoint model (reference: integrator.py:515-1100).

    Extra attributes as in the reference: ``tg_ic, n_tg_traj, func_jac``.
    """

    def __init__(self, num_threads=None, b=None, c=None, a=None, number_of_dimensions=None):
        _IntegratorBase.__init__(self, num_threads, b, c, a, number_of_dimensions)
        self.tg_ic = None
        self._recorded_fmatrix = None
        self.n_tg_traj = 0
        self._adjoint = False
        self._boundary = None
        self._inverse = 1.
        self.func_jac = None

    def set_func(self, f, fjac, ic_init=True):
        """Set the tendencies and Jacobian functions (integrator.py:659-684)."""
        if tensor_of(f) is not tensor_of(fjac, "fjac"):
            raise ValueError("f and fjac must come from the same create_tendencies() call")
        self.func = f
        self.func_jac = fjac
        if ic_init:
            self.ic = None
        self.start()

    def initialize(self, convergence_time, dt, pert_size=0.01, reconvergence_time=None, forward=True,
                   number_of_trajectories=1, ic=None, reconverge=False):
        """Converge to the attractor with the nonlinear model only (integrator.py:717-836)."""
        helper = RungeKuttaIntegrator(self.num_threads, self.b, self.c, self.a, self.n_dim)
        helper.set_func(self.func)
        helper.initialize(convergence_time, dt, pert_size, reconvergence_time, forward, number_of_trajectories, ic,
                          reconverge)
        self.ic = helper.ic

    def integrate(self, t0, t, dt, ic=None, tg_ic=None, forward=True, adjoint=False, inverse=False, boundary=None,
                  write_steps=1):
        """Integrate both systems from ``t0`` to ``t`` (integrator.py:838-1005); ``tg_ic`` shape rules are the
        reference's (integrator.py:933-958)."""
        if self.func is None or self.func_jac is None:
            print('No function to integrate defined!')
            return 0
        check_boundary(boundary)

        self._prepare(t0, t, dt, ic, forward, write_steps)

        if tg_ic is None:
            tg_ic = np.eye(self.ic.shape[1])

        tg_ic_sav = tg_ic.copy()
        self.tg_ic = normalise_tg_ic(tg_ic, self.n_traj, self.n_dim)
        self.n_tg_traj = self.tg_ic.shape[1]

        self._adjoint = adjoint
        self._boundary = _zeros_func if boundary is None else boundary
        self._inverse = 1.
        if inverse:
            self._inverse *= -1.

        self._recorded_traj, self._recorded_fmatrix = _integrate_runge_kutta_tgls_jit(
            self.func, self.func_jac, self._time, self.ic, self.tg_ic, self._time_direction, self._write_steps,
            self.b, self.c, self.a, self._adjoint, self._inverse, self._boundary)

        if len(tg_ic_sav.shape) == 2:
            if self._recorded_fmatrix.shape[1:3] != tg_ic_sav.shape:
                self._recorded_fmatrix = np.swapaxes(self._recorded_fmatrix, 1, 2)
        elif len(tg_ic_sav.shape) == 3:
            if tg_ic_sav.shape[1] != self.n_dim:
                if self._recorded_fmatrix.shape[:3] != tg_ic_sav.shape:
                    self._recorded_fmatrix = np.swapaxes(self._recorded_fmatrix, 1, 2)

    def get_trajectories(self):
        """``time, traj, tg_traj`` of the last integration (integrator.py:1007-1043)."""
        return self._returned_time(), np.squeeze(self._recorded_traj), np.squeeze(self._recorded_fmatrix)

    def get_tg_ic(self):
        """Returns the initial conditions of the linear ODEs stored in the integrator."""
        return self.tg_ic

    def set_tg_ic(self, tg_ic):
        """Direct setter for the integrator's linear ODEs initial conditions."""
        self.tg_ic = tg_ic


class TglsTrajectoryProcess(TrajectoryProcess):
    """Placeholder for the reference's TGLS worker class (integrator.py:1103-1169)."""
