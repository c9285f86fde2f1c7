"""Post-import adapters for reference modules that hand the tendencies to numba code.

``create_tendencies`` of the CUDA path returns callables that carry a device tensor handle; they are not numba
dispatchers.  One caller in the reference passes them into ``@njit`` code and therefore cannot take them as they are:
the vertical-velocity diagnostic, ``qgs/diagnostics/wind.py:678-679`` (construction) and ``:705-714``
(``_compute_omega_term``: a per-record loop ``func(time[i], data[:, i])`` over the whole trajectory).  When the overlay
package is active, importing that module swaps the loop for ONE batched evaluation on the device -- the records of a
trajectory are an ensemble of states as far as ``f`` is concerned.  Everything else in the module stays the
reference's code.  Like ``set_func``, the adapted function takes tensor-backed tendencies only.
"""
import importlib.abc
import sys

import numpy as np


def omega_term(time, data, func, thermo_func):
    """``f(t_i, x_i) - f_thermo(t_i, x_i)`` for every record ``x_i = data[:, i]`` (wind.py:705-714), both tendencies
    evaluated for all records in one launch each; the system is autonomous, ``time`` only fixes the record count."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim != 2 or data.shape[-1] != np.shape(time)[0]:
        raise ValueError("data must be (n_dim, n_records) with one record per time, got %s for %d times"
                         % (data.shape, np.shape(time)[0]))
    states = np.ascontiguousarray(data.T)
    return np.ascontiguousarray((func(0., states) - thermo_func(0., states)).T)


def _patch_wind(module):
    from qgs_b200.functions.tendencies import Tendencies

    def _compute_omega_term(time, data, func, thermo_func):
        if not (isinstance(func, Tendencies) and isinstance(thermo_func, Tendencies)):
            raise TypeError("the overlay evaluates the vertical-velocity term on the device: both functions must come "
                            "from create_tendencies / create_atmo_thermo_tendencies (there is no CPU fallback)")
        return omega_term(time, data, func, thermo_func)

    module._compute_omega_term = _compute_omega_term


PATCHES = {"qgs.diagnostics.wind": _patch_wind}


class _PatchingLoader(object):
    """The module's real loader plus one call after ``exec_module``; every other attribute is the real loader's."""

    def __init__(self, loader, patch):
        self._loader = loader
        self._patch = patch

    def create_module(self, spec):
        return self._loader.create_module(spec)

    def exec_module(self, module):
        self._loader.exec_module(module)
        self._patch(module)
        module._qgsb_patched = True

    def __getattr__(self, name):
        return getattr(self._loader, name)


class PatchAfterImport(importlib.abc.MetaPathFinder):
    """Meta-path entry that lets the regular finders locate a module of ``PATCHES`` and wraps its loader."""

    def find_spec(self, fullname, path=None, target=None):
        patch = PATCHES.get(fullname)
        if patch is None:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _PatchingLoader(spec.loader, patch)
                return spec
        return None


def install():
    """Idempotent; modules of ``PATCHES`` that were imported before are patched in place."""
    if not any(isinstance(f, PatchAfterImport) for f in sys.meta_path):
        sys.meta_path.insert(0, PatchAfterImport())
    for name, patch in PATCHES.items():
        module = sys.modules.get(name)
        if module is not None and not getattr(module, "_qgsb_patched", False):
            patch(module)
            module._qgsb_patched = True
