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


def _chunks(n_traj, num):
    """``num`` consecutive blocks of ``n_traj // num`` members; the last one also takes the remainder
    (statistics.py:38-62)."""
    size = n_traj // num
    edges = [k * size for k in range(num)] + [n_traj]
    return list(zip(edges[:-1], edges[1:]))


class TrajectoriesStatistics(object):
    """Same attributes and methods as the reference class: ``ic``, ``integrator``, ``func_list``, ``mean_func``;
    ``initialize``, ``set_func_list``, ``set_integrator``, ``set_ic``, ``compute_stats``, ``get_ic``, ``get_stats``."""

    def __init__(self):
        self.integrator = None
        self.ic = None
        self.func_list = []
        self.mean_func = []

    # ---- plumbing (statistics.py:16-31, 68-72) ------------------------------------------------------------
    def set_integrator(self, integrator):
        self.integrator = integrator

    def set_func_list(self, func_list):
        self.func_list = func_list

    def set_ic(self, ic):
        self.ic = ic

    def get_ic(self):
        return self.ic

    def get_stats(self):
        return self.mean_func

    def initialize(self, convergence_time, dt, pert_size=0.01, reconvergence_time=None, number_of_trajectories=1,
                   ic=None):
        """Spin the ensemble up with the integrator's own ``initialize`` and adopt its states."""
        self.integrator.initialize(convergence_time, dt, pert_size=pert_size, ic=ic,
                                   reconvergence_time=reconvergence_time,
                                   number_of_trajectories=number_of_trajectories)
        self.set_ic(self.integrator.get_ic())

    # ---- statistics ------------------------------------------------------------------------------------------
    def compute_stats(self, t0, t, dt, ic=None, forward=True, write_steps=1, num=1):
        """Ensemble mean of every function of ``func_list`` (host callables on ``(n_traj, n_dim, n_records)``
        arrays), computed block by block over ``num`` blocks of members and then averaged over the blocks --
        statistics.py:33-66, including its equal weighting of a larger last block."""
        if ic is not None:
            self.set_ic(ic)
        per_block = None
        for k, (lo, hi) in enumerate(_chunks(self.ic.shape[0], num)):
            self.integrator.integrate(t0, t, dt, ic=self.ic[lo:hi], forward=forward, write_steps=write_steps)
            traj = self.integrator.get_trajectories()[1]
            if per_block is None:
                per_block = np.zeros((len(self.func_list), num) + traj.shape[1:])
            for j, func in enumerate(self.func_list):
                per_block[j, k] = func(traj).mean(axis=0)
        self.mean_func = per_block.mean(axis=1)

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
