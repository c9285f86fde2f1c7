"""Quick device-side throughput probe (not the bench): FP64 peak, fused RK4 member-steps/s."""
import ctypes
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402

FLOPS = {"rp": 2620, "maooam36": 4132, "dynT": 5748, "T4": 103468, "atm6x6": 329488}


def run(name, N, steps, spec=True, reps=3):
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    n = int(z["ndim"])
    f, _ = tendencies_from_tensor(n, z["coo"], z["val"], z["jcoo"], z["jval"], specialise=spec)
    f.tensor.use_specialised(spec)
    lib = _lib.load()
    ens = ctypes.c_void_p()
    _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, N, ctypes.byref(ens)))
    ic = np.random.default_rng(1).random((N, n)) * 0.01
    _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic)))
    b, c, a = rk4_tableau()
    dt = np.full(steps, 0.1)
    ms = ctypes.c_double()
    best = 1e30
    for _ in range(reps):
        _lib.check(lib.qgsb_ensemble_integrate(ens, steps, _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c),
                                               ctypes.byref(ms)))
        best = min(best, ms.value)
    lib.qgsb_ensemble_destroy(ens)
    rate = N * steps / (best * 1e-3)
    print("%-9s kind=%d N=%8d steps=%5d  %9.3f ms  %.3e member-steps/s  %.2f TFLOP/s(alg)" %
          (name, f.tensor.kernel_kind, N, steps, best, rate, rate * FLOPS[name] / 1e12), flush=True)
    return rate


if __name__ == "__main__":
    _lib.init(0)
    print(_lib.device_info())
    print("fp64 peak (DFMA microbench): %.2f TFLOP/s" % _lib.fp64_peak(), flush=True)
    for N in (1 << 15, 1 << 17, 1 << 20):
        run("maooam36", N, 200)
    run("maooam36", 1 << 20, 1000)
    run("maooam36", 1 << 17, 200, spec=False)
    run("rp", 1 << 20, 200)
    run("dynT", 1 << 18, 200)
    run("dynT", 1 << 15, 100, spec=False)
    run("T4", 1 << 13, 20, spec=False)
    run("atm6x6", 1 << 12, 10, spec=False)
    print("fp64 peak again: %.2f TFLOP/s" % _lib.fp64_peak())
