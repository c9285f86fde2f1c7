"""Device-side throughput of the tangent-linear / Benettin kernels and of the generic RK kernels."""
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
from qgs_b200.toolbox.lyapunov import _subtimes  # noqa: E402


def load(name):
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    return tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])


def tgls(name, N, steps, m=None):
    f, Df = load(name)
    n = f.ndim
    m = m or n
    lib = _lib.load()
    b, c, a = rk4_tableau()
    rng = np.random.default_rng(0)
    ic = rng.random((N, n)) * 0.01
    tg = np.repeat(np.eye(n)[None, :, :m], N, axis=0).copy()
    dt = np.full(steps, 0.1)
    traj = np.empty((N, n, 1))
    fm = np.empty((N, n, m, 1))
    ms = ctypes.c_double()
    for _ in range(2):
        _lib.check(lib.qgsb_rk_tgls_integrate(f.tensor.handle, N, _lib.dptr(ic), m, _lib.dptr(tg), steps, _lib.dptr(dt),
                                              4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), 0, 1, 0, 1.0, 1,
                                              _lib.dptr(traj), _lib.dptr(fm), ctypes.byref(ms)))
    print("TGLS  %-9s N=%6d m=%3d steps=%4d %9.3f ms  %.3e member-steps/s" % (name, N, m, steps, ms.value,
                                                                             N * steps / ms.value * 1e3), flush=True)


def lyap(name, N, n_pre, n_rec, m=None, mdt=0.1):
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
    ws = 10
    R = len(tim[::ws]) + (1 if tim[::ws][-1] != tim[-1] else 0)
    rt, re, rv = np.empty((N, n, R)), np.empty((N, m, R)), np.empty((N, n, m, R))
    ms = ctypes.c_double()
    for _ in range(2):
        _lib.check(lib.qgsb_lyap_benettin(f.tensor.handle, N, _lib.dptr(ic), 0, m, _lib.dptr(q0), None, len(pre) - 1,
                                          len(tim) - 1, _lib.dptr(dtm), sub_ptr.ctypes.data_as(_lib.c_long_p),
                                          _lib.dptr(sub_dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ws, 0, 1.0, R,
                                          _lib.dptr(rt), _lib.dptr(re), _lib.dptr(rv), None, None, ctypes.byref(ms)))
    steps = n_pre + n_rec
    print("LYAP  %-9s N=%6d m=%3d steps=%4d %9.3f ms  %.3e member-steps/s" % (name, N, m, steps, ms.value,
                                                                             N * steps / ms.value * 1e3), flush=True)


def cpu_lyap(name, N, n_pre, n_rec):
    import oracle
    T = oracle.Tensor.from_npz(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    n = T.ndim
    b, c, a = oracle.rk4_tableau()
    rng = np.random.default_rng(0)
    ic = rng.random((N, n)) * 0.01
    q0 = np.stack([np.linalg.qr(rng.random((n, n)))[0] for _ in range(N)])
    pre = np.concatenate((np.arange(0., n_pre * 0.1, 0.1), [n_pre * 0.1]))
    tim = np.concatenate((np.arange(n_pre * 0.1, (n_pre + n_rec) * 0.1, 0.1), [(n_pre + n_rec) * 0.1]))
    t0 = time.perf_counter()
    oracle.compute_backward_lyap(T, pre, tim, 0.1, ic, n, 10, False, 1., b, c, a, q0, None)
    w = time.perf_counter() - t0
    print("CPU LYAP %-9s N=%d steps=%d  %.2f s  %.3e member-steps/s (%d threads)" %
          (name, N, n_pre + n_rec, w, N * (n_pre + n_rec) / w, oracle.num_threads()), flush=True)


def main():
    _lib.init(0)
    tgls("maooam36", 8192, 50)
    tgls("maooam36", 8192, 50, m=1)
    tgls("rp", 8192, 50)
    lyap("maooam36", 8192, 20, 80)
    lyap("maooam36", 8192, 20, 80, m=10)
    lyap("maooam36", 2048, 10, 40, mdt=0.02)
    lyap("rp", 8192, 20, 80)
    lyap("dynT", 2048, 10, 40)
    tgls("T4", 512, 5)
    tgls("atm6x6", 256, 3, m=16)
    cpu_lyap("maooam36", 64, 10, 40)
    print("fp64 DFMA peak %.2f TFLOP/s, DMMA (mma.sync m8n8k4) peak %.2f TFLOP/s" % (_lib.fp64_peak(), _lib.dmma_peak()))


if __name__ == "__main__":
    main()
