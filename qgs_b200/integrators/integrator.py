"""Ensemble integrator classes on the GPU -- mirror of ``qgs/integrators/integrator.py``.

``RungeKuttaIntegrator`` and ``RungeKuttaTglsIntegrator`` keep the reference's constructor,
methods, attributes, shape conventions and quirks (integrator.py:27-450, :515-1100).  What is gone is
the pool of ``multiprocessing`` workers fed one pickled message per trajectory
(integrator.py:121-142, 388-395): ``integrate`` hands the whole ensemble to one fused CUDA launch.
``num_threads`` is kept for signature parity; it only sets the batch size of ``initialize`` like in
the reference (integrator.py:257-291).
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

    def initialize(self, convergence_time, dt, pert_size=0.01, reconvergence_time=None, forward=True,
                   number_of_trajectories=1, ic=None, reconverge=False):
        """Converge to the attractor by integrating over a transient (integrator.py:198-295).  Semantics are the
        reference's, including batches of ``num_threads`` members when reconverging."""
        if reconverge is None:
            reconverge = False

        if ic is None:
            i = self._infer_ndim()

            if number_of_trajectories > self.num_threads:
                reconverge = True
                tmp_ic = np.zeros((number_of_trajectories, i))
                tmp_ic[:self.num_threads] = np.random.randn(self.num_threads, i)
            else:
                tmp_ic = np.random.randn(number_of_trajectories, i)
        else:
            tmp_ic = ic.copy()
            if len(tmp_ic.shape) > 1:
                number_of_trajectories = tmp_ic.shape[0]

        if reconverge and reconvergence_time is not None:
            self.integrate(0., convergence_time, dt, ic=tmp_ic[:self.num_threads], write_steps=0, forward=forward)
            t, x = self.get_trajectories()
            x = np.atleast_2d(x)
            tmp_ic[:self.num_threads] = x
            if number_of_trajectories - self.num_threads > self.num_threads:
                next_len = self.num_threads
            else:
                next_len = number_of_trajectories - self.num_threads

            index = self.num_threads
            while True:
                perturbation = pert_size * np.random.randn(next_len, x.shape[1])
                self.integrate(0., reconvergence_time, dt, ic=x[:next_len] + perturbation, write_steps=0,
                               forward=forward)
                t, x = self.get_trajectories()
                x = np.atleast_2d(x)
                tmp_ic[index:index + next_len] = x
                index += next_len
                if number_of_trajectories - index > self.num_threads:
                    next_len = self.num_threads
                else:
                    next_len = number_of_trajectories - index
                if next_len <= 0:
                    break
            self.ic = tmp_ic
        else:
            self.integrate(0., convergence_time, dt, ic=tmp_ic, write_steps=0, forward=forward)
            t, x = self.get_trajectories()
            self.ic = x

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
