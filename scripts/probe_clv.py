"""Covariant Lyapunov vectors end to end through the reference-facing class (both methods), MAOOAM-36."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.toolbox.lyapunov import CovariantLyapunovsEstimator  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.init(0)
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
f, Df = tendencies_from_tensor(36, z["coo"], z["val"], z["jcoo"], z["jval"])
for method in (0, 1):
    for N in (1, 512):
        ic = np.random.default_rng(0).random((N, 36)) * 0.01
        est = CovariantLyapunovsEstimator(method=method)
        est.set_func(f, Df)
        est.compute_clvs(0., 1., 2., 3., 0.1, 0.1, ic=ic, write_steps=10, method=method)      # warm-up
        t0 = time.perf_counter()
        est.compute_clvs(0., 10., 30., 40., 0.1, 0.1, ic=ic, write_steps=10, method=method)
        w = time.perf_counter() - t0
        steps = 400
        print("CLV method %d  N=%4d  %d steps (t0..tc)  %.3f s  %.3e member-steps/s" % (method, N, steps, w, N * steps / w),
              flush=True)
