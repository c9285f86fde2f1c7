"""Host-side latency of single calls: the raw contractions (tensor passed as host arrays every time) and the f / Df
closures on one state."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions import sparse_mul as sm  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402

_lib.init(0)
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
x = np.random.default_rng(0).random(37)
x[0] = 1.


def timeit(label, fn, reps=2000):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    print("%-44s %7.1f us per call" % (label, (time.perf_counter() - t0) / reps * 1e6), flush=True)


timeit("sparse_mul3(coo, val, x, x)  MAOOAM-36", lambda: sm.sparse_mul3(z["coo"], z["val"], x, x))
timeit("sparse_mul2(jcoo, jval, x)   MAOOAM-36", lambda: sm.sparse_mul2(z["jcoo"], z["jval"], x))
coo, val, jcoo, jval = z["coo"], z["val"], z["jcoo"], z["jval"]
timeit("  ... arrays taken out of the npz once", lambda: sm.sparse_mul3(coo, val, x, x))
timeit("f(t, x) on one state", lambda: f(0., x[1:]))
timeit("Df(t, x) on one state", lambda: Df(0., x[1:]))
