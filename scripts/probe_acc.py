"""A/B of the MAOOAM-36 fused RK4 kernel: weighted stage sum in registers (2 blocks/SM) vs in shared memory (3 blocks/SM)."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402

_lib.init(0)
lib = _lib.load()
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
f, _ = tendencies_from_tensor(36, z["coo"], z["val"], z["jcoo"], z["jval"])
N, steps = 1 << 20, 500
ic = np.random.default_rng(1).random((N, 36)) * 0.01
b, c, a = rk4_tableau()
dt = np.full(steps, 0.1)


def run(tag):
    ens = ctypes.c_void_p()
    _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, N, ctypes.byref(ens)))
    _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic)))
    ms = ctypes.c_double()
    best = 1e30
    for _ in range(4):
        _lib.check(lib.qgsb_ensemble_integrate(ens, steps, _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c),
                                               ctypes.byref(ms)))
        best = min(best, ms.value)
    out = np.empty((N, 36))
    _lib.check(lib.qgsb_ensemble_download(ens, _lib.dptr(out)))
    lib.qgsb_ensemble_destroy(ens)
    print("%-28s %9.3f ms  %.4e member-steps/s  %.2f TFLOP/s  checksum %.15e" %
          (tag, best, N * steps / best * 1e3, N * steps / best * 1e3 * 4132 / 1e12, out.sum()), flush=True)


run("linked module (registers)")
for variant in sys.argv[1:]:
    _lib.check(lib.qgsb_load_plugin(os.path.join(REPO, "qgs_b200", "_jit", "exp", variant).encode()))
    f.tensor.use_specialised(True)
    run(variant)
print("fp64 peak %.2f" % _lib.fp64_peak())
