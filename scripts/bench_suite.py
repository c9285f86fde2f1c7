"""Secondary measurements for every configuration of BASELINE.json (bench.py carries the headline only).

    python scripts/bench_suite.py [out.json]

Device-timed (CUDA events inside libqgsb) member-steps/s of the fused RK4 kernels for the five tensors, of the
tangent-linear and Benettin kernels for MAOOAM-36, and the algorithmic FP64 rate of each against the DFMA peak
measured in the same process.  Flop counts per member-step follow SURVEY.md section 8(d).
"""
import ctypes
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrate import rk4_tableau  # noqa: E402
from qgs_b200.toolbox.lyapunov import _subtimes  # noqa: E402

FLOPS_RK = {"rp": 2620, "maooam36": 4132, "dynT": 5748, "T4": 103468, "atm6x6": 329488}


def load(name, spec=True):
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"], specialise=spec)
    return f, Df, z


def lyap_flops(z, m):
    """4 (F_f + 2 jnnz + 2 n^2 m) + 14 n + 14 n m + 4 n m^2 - 4/3 m^3 (dense-equivalent)."""
    n = int(z["ndim"])
    coo = z["coo"]
    f_f = int(sum(1 + np.count_nonzero(c[1:]) for c in coo if c[0] > 0))
    return 4 * (f_f + 2 * len(z["jval"]) + 2 * n * n * m) + 14 * n + 14 * n * m + 4 * n * m * m - 4. * m ** 3 / 3.


def tgls_flops(z, m):
    n = int(z["ndim"])
    coo = z["coo"]
    f_f = int(sum(1 + np.count_nonzero(c[1:]) for c in coo if c[0] > 0))
    return 4 * (f_f + 2 * len(z["jval"]) + 2 * n * n * m) + 14 * n + 14 * n * m


def rk(name, N, steps, ws=0):
    f, _, z = load(name)
    lib = _lib.load()
    ens = ctypes.c_void_p()
    _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, N, ctypes.byref(ens)))
    scale = 0.1 if name in ("rp", "tlad") else 0.01
    ic = np.random.default_rng(1).random((N, f.ndim)) * scale
    _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic)))
    b, c, a = rk4_tableau()
    dt = np.full(steps, 0.1)
    ms = ctypes.c_double()
    best = 1e30
    for _ in range(3):
        _lib.check(lib.qgsb_ensemble_integrate(ens, steps, _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c),
                                               ctypes.byref(ms)))
        best = min(best, ms.value)
    lib.qgsb_ensemble_destroy(ens)
    rate = N * steps / best * 1e3
    return {"what": "rk4", "tensor": name, "n_dim": f.ndim, "kernel_kind": f.tensor.kernel_kind, "members": N,
            "steps": steps, "ms": best, "member_steps_per_s": rate, "flops_per_member_step": FLOPS_RK[name],
            "tflops_algorithmic": rate * FLOPS_RK[name] / 1e12}


def tangent(name, N, steps, m, lyap):
    f, Df, z = load(name)
    n = f.ndim
    lib = _lib.load()
    b, c, a = rk4_tableau()
    rng = np.random.default_rng(0)
    ic = rng.random((N, n)) * (0.1 if name == "rp" else 0.01)
    ms = ctypes.c_double()
    if not lyap:
        tg = np.repeat(np.eye(n)[None, :, :m], N, axis=0).copy()
        dt = np.full(steps, 0.1)
        traj, fm = np.empty((N, n, 1)), np.empty((N, n, m, 1))
        for _ in range(2):
            _lib.check(lib.qgsb_rk_tgls_integrate(f.tensor.handle, N, _lib.dptr(ic), m, _lib.dptr(tg), steps,
                                                  _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), 0, 1, 0,
                                                  1.0, 1, _lib.dptr(traj), _lib.dptr(fm), ctypes.byref(ms)))
        flops = tgls_flops(z, m)
    else:
        n_pre, n_rec = steps // 5, steps - steps // 5
        q0 = np.stack([np.linalg.qr(rng.random((n, m)))[0] for _ in range(16)])
        q0 = np.ascontiguousarray(np.tile(q0, (N // 16 + 1, 1, 1))[:N])
        pre = np.concatenate((np.arange(0., n_pre * 0.1, 0.1), [n_pre * 0.1]))[:n_pre + 1]
        tim = (n_pre * 0.1 + np.concatenate((np.arange(0., n_rec * 0.1, 0.1), [n_rec * 0.1])))[:n_rec + 1]
        pa, sa = _subtimes(pre, 0.1)
        pb, sb = _subtimes(tim, 0.1)
        sub_ptr = np.ascontiguousarray(np.concatenate((pa, pb[1:] + pa[-1])), dtype=np.int64)
        sub_dt = np.ascontiguousarray(np.concatenate((sa, sb)))
        dtm = np.ascontiguousarray(np.concatenate((np.diff(pre), np.diff(tim))))
        ws = 10
        R = len(tim[::ws]) + (1 if tim[::ws][-1] != tim[-1] else 0)
        rt, re, rv = np.empty((N, n, R)), np.empty((N, m, R)), np.empty((N, n, m, R))
        for _ in range(2):
            _lib.check(lib.qgsb_lyap_benettin(f.tensor.handle, N, _lib.dptr(ic), 0, m, _lib.dptr(q0), None,
                                              len(pre) - 1, len(tim) - 1, _lib.dptr(dtm),
                                              sub_ptr.ctypes.data_as(_lib.c_long_p), _lib.dptr(sub_dt), 4,
                                              _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), ws, 0, 1.0, R, _lib.dptr(rt),
                                              _lib.dptr(re), _lib.dptr(rv), None, None, ctypes.byref(ms)))
        steps = len(pre) + len(tim) - 2
        flops = lyap_flops(z, m)
    rate = N * steps / ms.value * 1e3
    return {"what": "benettin" if lyap else "tgls", "tensor": name, "n_dim": n, "n_vec": m, "members": N,
            "steps": steps, "ms": ms.value, "member_steps_per_s": rate, "flops_per_member_step": flops,
            "tflops_algorithmic": rate * flops / 1e12}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    _lib.init(0)
    peak = _lib.fp64_peak()
    rows = [rk("maooam36", 1 << 20, 1000), rk("rp", 1 << 20, 500), rk("dynT", 1 << 20, 200),
            rk("T4", 148 * 2 * 128 * 4, 50), rk("atm6x6", 148 * 96 * 4, 20),
            tangent("maooam36", 8192, 50, 36, False), tangent("maooam36", 8192, 100, 36, True),
            tangent("maooam36", 8192, 100, 10, True), tangent("rp", 8192, 100, 20, True),
            tangent("dynT", 4096, 50, 38, True)]
    for r in rows:
        r["fp64_peak_tflops"] = peak
        r["frac_of_fp64_peak"] = r["tflops_algorithmic"] / peak
        print(json.dumps(r), flush=True)
    if out:
        with open(out, "w") as fh:
            json.dump({"fp64_peak_tflops_measured": peak, "device": _lib.device_info(), "rows": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
