"""Latency regime of the Benettin kernels: few trajectories, long windows (the notebooks' usage)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.toolbox.lyapunov import LyapunovsEstimator  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.init(0)
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
f, Df = tendencies_from_tensor(36, z["coo"], z["val"], z["jcoo"], z["jval"])
est = LyapunovsEstimator()
est.set_func(f, Df)
for N in (1, 7, 64):
    ic = np.random.default_rng(0).random((N, 36)) * 0.01
    for n_vec in (36, 10):
        est.compute_lyapunovs(0., 1., 2., 0.1, 0.1, ic=ic, write_steps=10, n_vec=n_vec)
        steps = 5000
        t0 = time.perf_counter()
        est.compute_lyapunovs(0., 100., steps * 0.1, 0.1, 0.1, ic=ic, write_steps=10, n_vec=n_vec)
        w = time.perf_counter() - t0
        print("LYAP maooam36 N=%3d n_vec=%2d steps=%d  %.3f s  %.1f us/step  %.3e member-steps/s" %
              (N, n_vec, steps, w, w / steps * 1e6, N * steps / w), flush=True)
