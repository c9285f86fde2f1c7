"""Cholesky QR on the unobserved steps of the Benettin loop (QGSB_QR_CHOL, read at every launch) against Householder
everywhere: device time and the largest difference of the recorded exponents / vectors."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402
from qgs_b200.toolbox.lyapunov import _subtimes  # noqa: E402
from scripts.perf_probe2 import load  # noqa: E402


def run(name, N, n_pre, n_rec, m=None, vectors=True, ws=10, mdt=0.1):
    f, Df = load(name)
    n = f.ndim
    m = m or n
    lib = _lib.load()
    b, c, a = rk4_tableau()
    rng = np.random.default_rng(0)
    ic = rng.random((N, n)) * 0.01
    q0 = np.stack([np.linalg.qr(rng.random((n, m)))[0] for _ in range(min(N, 16))])
    q0 = np.ascontiguousarray(np.tile(q0, (N // len(q0) + 1, 1, 1))[:N])
    pre = np.concatenate((np.arange(0., n_pre * 0.1, 0.1), [n_pre * 0.1]))
    tim = np.concatenate((np.arange(n_pre * 0.1, (n_pre + n_rec) * 0.1, 0.1), [(n_pre + n_rec) * 0.1]))
    pa, sa = _subtimes(pre, mdt)
    pb, sb = _subtimes(tim, mdt)
    sub_ptr = np.ascontiguousarray(np.concatenate((pa, pb[1:] + pa[-1])), dtype=np.int64)
    sub_dt = np.ascontiguousarray(np.concatenate((sa, sb)))
    dtm = np.ascontiguousarray(np.concatenate((np.diff(pre), np.diff(tim))))
    R = len(tim[::ws]) + (1 if tim[::ws][-1] != tim[-1] else 0)
    out = {}
    for mode in ("0", "1"):
        os.environ["QGSB_QR_CHOL"] = mode
        rt, re = np.empty((N, n, R)), np.empty((N, m, R))
        rv = np.empty((N, n, m, R)) if vectors else None
        ms = ctypes.c_double()
        for _ in range(2):
            _lib.check(lib.qgsb_lyap_benettin(f.tensor.handle, N, _lib.dptr(ic), 0, m, _lib.dptr(q0), None, len(pre) - 1,
                                              len(tim) - 1, _lib.dptr(dtm), sub_ptr.ctypes.data_as(_lib.c_long_p),
                                              _lib.dptr(sub_dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ws, 0, 1.0, R,
                                              _lib.dptr(rt), _lib.dptr(re), _lib.dptr(rv), None, None, ctypes.byref(ms)))
        out[mode] = (ms.value, rt, re, rv)
    steps = n_pre + n_rec
    d_exp = np.abs(out["0"][2] - out["1"][2]).max()
    d_vec = np.abs(out["0"][3] - out["1"][3]).max() if vectors else float("nan")
    print("%-9s N=%6d m=%3d steps=%4d vectors=%d ws=%2d | householder %8.3f ms %.3e | cholesky %8.3f ms %.3e | x%.2f | "
          "max|d exp| %.2e  max|d vec| %.2e  max|exp| %.2f" %
          (name, N, m, steps, vectors, ws, out["0"][0], N * steps / out["0"][0] * 1e3, out["1"][0],
           N * steps / out["1"][0] * 1e3, out["0"][0] / out["1"][0], d_exp, d_vec, np.abs(out["0"][2]).max()), flush=True)


if __name__ == "__main__":
    _lib.init(0)
    run("maooam36", 8192, 20, 80, vectors=False)
    run("maooam36", 8192, 20, 80, vectors=True)
    run("maooam36", 8192, 20, 80, m=10, vectors=False)
    run("maooam36", 8192, 20, 80, m=20, vectors=False)
    run("rp", 8192, 20, 80, vectors=False)
    run("rp", 8192, 20, 80, vectors=True, ws=1)
    run("dynT", 2048, 10, 40, vectors=False)
    run("maooam36", 2048, 10, 40, m=5, vectors=True, mdt=0.02)
