"""Ensemble statistics -- mirror of ``qgs/integrators/statistics.py`` (``TrajectoriesStatistics``, :7-75).

The reference integrates the ensemble in ``num`` chunks, pulls every ``(n_traj, n_dim, n_records)`` trajectory block
back from its worker pool and averages user functions of it over the members.  The class below keeps that interface
(same methods, same shapes of ``get_stats()``); arbitrary Python functions of the trajectories still see host arrays
because they are host code.  For the functions that do not need the trajectories -- the first two moments --
:meth:`compute_moments` runs on the device-resident ensemble (``qgsb_ensemble_integrate_moments``): records are
reduced in HBM, the members may be sharded over ranks and only ``2 * n_records * n_dim`` sums are exchanged.
"""
import numpy as np

from qgs_b200.ensemble import DeviceEnsemble


class TrajectoriesStatistics(object):

    def __init__(self):
        self.ic = None
        self.integrator = None
        self.func_list = list()
        self.mean_func = list()

    def initialize(self, convergence_time, dt, pert_size=0.01, reconvergence_time=None,
                   number_of_trajectories=1, ic=None):
        """statistics.py:16-21"""
        self.integrator.initialize(convergence_time, dt, pert_size=pert_size, reconvergence_time=reconvergence_time,
                                   number_of_trajectories=number_of_trajectories, ic=ic)
        self.ic = self.integrator.get_ic()

    def set_func_list(self, func_list):
        self.func_list = func_list

    def set_integrator(self, integrator):
        self.integrator = integrator

    def set_ic(self, ic):
        self.ic = ic

    def compute_stats(self, t0, t, dt, ic=None, forward=True, write_steps=1, num=1):
        """statistics.py:33-66: ensemble mean of every function of ``func_list`` over ``num`` chunks of the initial
        conditions (chunk ``num - 1`` takes the remainder), then the mean over the chunks."""
        if ic is not None:
            self.set_ic(ic)
        number_of_trajectories = self.ic.shape[0]
        sub = number_of_trajectories // num
        bounds = [(i * sub, (i + 1) * sub) for i in range(num - 1)] + [((num - 1) * sub, number_of_trajectories)]
        realization = None
        for i, (lo, hi) in enumerate(bounds):
            self.integrator.integrate(t0, t, dt, ic=self.ic[lo:hi], forward=forward, write_steps=write_steps)
            time, traj = self.integrator.get_trajectories()
            if realization is None:
                realization = np.zeros((len(self.func_list), num, traj.shape[1], traj.shape[2]))
            for j, f in enumerate(self.func_list):
                realization[j, i] = np.mean(f(traj), axis=0)
        self.mean_func = np.mean(realization, axis=1)

    def compute_moments(self, t0, t, dt, ic=None, forward=True, write_steps=1, sharded=False):
        """Mean and variance of every variable at every record, computed on the device without the trajectory
        dump.  Returns ``(time, mean, var)`` with ``mean``, ``var`` of shape ``(n_dim, n_records)`` like the
        entries of ``get_stats()``; also stores ``[mean, mean_of_squares]`` as ``get_stats()`` would for
        ``func_list = [lambda x: x, lambda x: x**2]`` with ``num=1``."""
        if ic is not None:
            self.set_ic(ic)
        ens = DeviceEnsemble(self.integrator.func, self.ic, sharded=sharded)
        try:
            time, mean, var = ens.integrate_moments(t0, t, dt, forward=forward, write_steps=write_steps,
                                                    b=self.integrator.b, c=self.integrator.c, a=self.integrator.a)
        finally:
            ens.close()
        mean, var = mean.T, var.T
        self.mean_func = np.stack((mean, var + mean * mean))
        return time, mean, var

    def get_ic(self):
        return self.ic

    def get_stats(self):
        return self.mean_func
