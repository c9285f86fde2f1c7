"""Latency regime: few trajectories, many steps (the reference's example scripts integrate ONE trajectory)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrator import RungeKuttaIntegrator  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.init(0)
for name in sys.argv[1:] or ["maooam36"]:
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
    integ = RungeKuttaIntegrator()
    integ.set_func(f)
    for N in (1, 8, 64, 512):
        ic = np.random.default_rng(0).random((N, f.ndim)) * 0.01
        for steps, ws in ((20000, 0), (20000, 10)):
            integ.integrate(0., 10., 0.1, ic=ic, write_steps=ws)     # warm-up
            t0 = time.perf_counter()
            integ.integrate(0., steps * 0.1, 0.1, ic=ic, write_steps=ws)
            integ.get_trajectories()
            w = time.perf_counter() - t0
            print("%-9s N=%4d steps=%6d ws=%2d  %.3f s  %.2f us/step  %.3e member-steps/s" %
                  (name, N, steps, ws, w, w / steps * 1e6, N * steps / w), flush=True)
