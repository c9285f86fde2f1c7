"""T4 / dynT specialised kernels and the trajectory-recording regime of the MAOOAM-36 kernel."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402
from scripts.perf_probe import run  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def recording(N, steps, ws):
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
    f, _ = tendencies_from_tensor(36, z["coo"], z["val"], z["jcoo"], z["jval"])
    lib = _lib.load()
    b, c, a = rk4_tableau()
    ic = np.random.default_rng(1).random((N, 36)) * 0.01
    dt = np.full(steps, 0.1)
    L = steps + 1
    R = 1 if ws == 0 else (L + ws - 1) // ws + (1 if ((L + ws - 1) // ws - 1) * ws != L - 1 else 0)
    out = np.empty((N, 36, R))
    ms = ctypes.c_double()
    for _ in range(2):
        _lib.check(lib.qgsb_rk_integrate(f.tensor.handle, N, _lib.dptr(ic), steps, _lib.dptr(dt), 4, _lib.dptr(a),
                                         _lib.dptr(b), _lib.dptr(c), ws, 1, R, _lib.dptr(out), ctypes.byref(ms)))
    rate = N * steps / ms.value * 1e3
    print("REC maooam36 N=%d steps=%d ws=%d R=%d  %.3f ms  %.3e member-steps/s  record writes %.1f GB/s (device part)"
          % (N, steps, ws, R, ms.value, rate, N * 36 * 8 * R / ms.value / 1e6), flush=True)


if __name__ == "__main__":
    _lib.init(0)
    run("T4", 1 << 17, 50)
    run("T4", 1 << 13, 20, spec=False)
    run("dynT", 1 << 20, 200)
    recording(1 << 18, 1000, 10)
    recording(1 << 18, 200, 1)
    recording(1 << 20, 100, 1)
