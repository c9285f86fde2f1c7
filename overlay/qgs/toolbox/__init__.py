import os as _os

from qgs import _real_package_dirs

__path__ = [_os.path.dirname(_os.path.abspath(__file__))] + _real_package_dirs("toolbox")
