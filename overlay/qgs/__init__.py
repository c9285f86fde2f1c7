"""Drop-in overlay: put this directory BEFORE the reference on ``sys.path`` / ``PYTHONPATH``.

``import qgs`` then resolves to this package, which (i) serves the five hot-path modules

    qgs.functions.sparse_mul, qgs.functions.tendencies, qgs.integrators.integrate,
    qgs.integrators.integrator, qgs.toolbox.lyapunov

from ``qgs_b200`` (CUDA path) and (ii) extends its ``__path__`` with the real qgs package so that
every other sub-package (params, basis, inner_products, tensors, diagnostics, ...) is the untouched
reference.  User scripts such as ``qgs_maooam.py`` run unchanged.  The one reference module that passes the tendencies
into ``@njit`` code (``qgs.diagnostics.wind``, vertical velocity) is adapted on import, see ``qgs_b200/overlay_hooks.py``.

The real package is looked up in ``$QGS_REFERENCE_PATH`` (the directory that contains ``qgs/``) or
further down ``sys.path``.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
if _REPO not in sys.path:
    sys.path.append(_REPO)          # makes qgs_b200 importable


def _real_package_dirs(subpackage=""):
    found = []
    candidates = []
    if os.environ.get("QGS_REFERENCE_PATH"):
        candidates.append(os.environ["QGS_REFERENCE_PATH"])
    candidates += [p for p in sys.path if p]
    for base in candidates:
        d = os.path.join(base, "qgs", subpackage) if subpackage else os.path.join(base, "qgs")
        if os.path.isdir(d) and os.path.abspath(d) != os.path.abspath(os.path.join(_HERE, subpackage)) \
                and os.path.exists(os.path.join(d, "__init__.py")) and d not in found:
            found.append(d)
    return found


__path__ = [_HERE] + _real_package_dirs()


def _real_version():
    """``__version__`` of the real package (its __init__ is shadowed by this one): qgs/__init__.py:2."""
    import re
    for d in _real_package_dirs():
        try:
            with open(os.path.join(d, "__init__.py")) as fh:
                m = re.search(r"__version__\s*=\s*['\"]([^'\"]+)['\"]", fh.read())
        except OSError:
            continue
        if m:
            return m.group(1)
    return ""


__version__ = _real_version()

# reference modules that pass the tendencies into numba code get a batched device evaluation instead (wind.py:705-714)
try:
    from qgs_b200 import overlay_hooks as _hooks
    _hooks.install()
except ImportError:
    pass

# tensor construction in the reference imports pydata `sparse` and `pebble`; use the stand-ins when absent
try:
    from qgs_b200 import compat as _compat
    _compat.install()
except ImportError:
    pass
