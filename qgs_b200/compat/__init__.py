"""Stand-ins for two third-party packages the reference imports at tensor-construction time.

The reference builds its tendencies tensor with pydata ``sparse`` (unpinned:
``/root/reference/requirements.txt:11``) and fans symbolic inner products over a
``pebble.ProcessPool`` (``qgs/inner_products/symbolic.py:26``).  Neither package is present in the
build image, and tensor construction is out of the hot-path scope,
so this sub-package ships the small subset of both APIs that the reference touches.  They are only
used to *produce* the ``coo, val`` arrays handed to the CUDA path; no arithmetic of the hot path
runs through them.

``install()`` puts the stand-ins on ``sys.path`` only when the real packages are missing.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install(force=False):
    """Make ``import sparse`` / ``import pebble`` work.  Real packages win unless ``force``."""
    need = force or importlib.util.find_spec("sparse") is None or importlib.util.find_spec("pebble") is None
    if need and _HERE not in sys.path:
        sys.path.append(_HERE) if not force else sys.path.insert(0, _HERE)
    return need
