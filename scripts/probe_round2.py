"""Round-2 secondary measurements: non-chain tableaux on the tensor-specialised kernel (Kutta's 3/8 rule and third-order
rule against classic RK4, MAOOAM-36, 2^20 members) and the tangent-linear / Benettin kernels at the 228 variables of the
6x6 model (generic kernels: one block per member, matrices in a global scratch)."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from scripts import bench_suite as bs  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402

_lib.init(0)
lib = _lib.load()
TABLEAUX = {
    "rk4": rk4_tableau(),
    "kutta38": (np.array([1., 3., 3., 1.]) / 8., np.array([0., 1. / 3, 2. / 3, 1.]),
                np.array([[0., 0., 0., 0.], [1. / 3, 0., 0., 0.], [-1. / 3, 1., 0., 0.], [1., -1., 1., 0.]])),
    "kutta3": (np.array([1., 4., 1.]) / 6., np.array([0., .5, 1.]),
               np.array([[0., 0., 0.], [.5, 0., 0.], [-1., 2., 0.]])),
}


def rk(name, tableau, N, steps, spec=True):
    f, _, z = bs.load(name, spec=spec)
    ens = ctypes.c_void_p()
    _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, N, ctypes.byref(ens)))
    ic = np.random.default_rng(1).random((N, f.ndim)) * 0.01
    _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic)))
    b, c, a = (_lib.f64(v) for v in TABLEAUX[tableau])
    dt = np.full(steps, 0.1)
    ms = ctypes.c_double()
    best = 1e30
    for _ in range(3):
        _lib.check(lib.qgsb_ensemble_integrate(ens, steps, _lib.dptr(dt), len(b), _lib.dptr(a), _lib.dptr(b),
                                               _lib.dptr(c), ctypes.byref(ms)))
        best = min(best, ms.value)
    lib.qgsb_ensemble_destroy(ens)
    print("RK %-8s %-8s spec=%d N=%d steps=%d  %8.2f ms  %.4g member-steps/s" % (name, tableau, spec, N, steps, best,
                                                                                  N * steps / best * 1e3), flush=True)


rk("maooam36", "rk4", 1 << 20, 200)
rk("maooam36", "kutta38", 1 << 20, 200)
rk("maooam36", "kutta3", 1 << 20, 200)
rk("maooam36", "kutta38", 1 << 17, 50, spec=False)
rk("dynT", "kutta38", 1 << 20, 100)
for m in (16, 228):
    r = bs.tangent("atm6x6", 512, 4, m, False)
    print("TGLS atm6x6 m=%3d  %8.2f ms  %.4g member-steps/s  %.3g TFLOP/s dense-equivalent" %
          (m, r["ms"], r["member_steps_per_s"], r["tflops_algorithmic"]), flush=True)
for m in (16, 228):
    t0 = time.time()
    r = bs.tangent("atm6x6", 256, 5, m, True)
    print("BENETTIN atm6x6 m=%3d  %8.2f ms  %.4g member-steps/s  %.3g TFLOP/s dense-equivalent" %
          (m, r["ms"], r["member_steps_per_s"], r["tflops_algorithmic"]), flush=True)
