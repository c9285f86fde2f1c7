"""One process, several GPUs, the UNCHANGED reference-facing API: RungeKuttaIntegrator.integrate() and
LyapunovsEstimator.compute_lyapunovs() on one device and on every visible device -- wall-clock end to end (host numpy
arrays in and out) and bitwise comparison of the results.

    python scripts/multi_device_api.py [log2_members] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrator import RungeKuttaIntegrator  # noqa: E402
from qgs_b200.toolbox.lyapunov import LyapunovsEstimator  # noqa: E402


def main():
    import torch
    G = torch.cuda.device_count()
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20 + max(0, G.bit_length() - 1)
    out = sys.argv[2] if len(sys.argv) > 2 else None
    N = 1 << log2n
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
    _lib.set_devices([0])
    f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
    ic = np.random.default_rng(21217).random((N, 36)) * 0.01
    Nl = 8192 * G
    res = {"devices": G, "members": N, "lyapunov_members": Nl}
    ends, lyaps = {}, {}
    for devs in ([0], list(range(G))):
        _lib.set_devices(devs)
        integ = RungeKuttaIntegrator()
        integ.set_func(f)
        integ.integrate(0., 10., 0.1, ic=ic[:N // 4], write_steps=0)        # warm-up: contexts, replicas, pools
        t0 = time.perf_counter()
        integ.integrate(0., 100., 0.1, ic=ic, write_steps=0)
        _, end = integ.get_trajectories()
        wall = time.perf_counter() - t0
        ends[len(devs)] = end
        res["rk4_e2e_%d" % len(devs)] = N * 1000 / wall
        est = LyapunovsEstimator()
        est.set_func(f, Df)
        np.random.seed(1)
        est.compute_lyapunovs(0., 1., 2., 0.1, 0.1, ic[:Nl], write_steps=10, vectors=False)     # warm-up
        np.random.seed(1)
        t0 = time.perf_counter()
        est.compute_lyapunovs(0., 20., 40., 0.1, 0.1, ic[:Nl], write_steps=10, vectors=False)
        ly = est.get_lyapunovs()
        wall = time.perf_counter() - t0
        lyaps[len(devs)] = ly
        res["lyapunov_e2e_%d" % len(devs)] = Nl * 400 / wall
        print(devs, "rk4 %.4g member-steps/s, lyapunov %.4g" % (res["rk4_e2e_%d" % len(devs)],
                                                                 res["lyapunov_e2e_%d" % len(devs)]), flush=True)
    _lib.set_devices([0])
    if G > 1:
        res["rk4_bitwise_equal"] = bool(np.array_equal(ends[1], ends[G]))
        res["lyapunov_bitwise_equal"] = bool(np.array_equal(lyaps[1][1], lyaps[G][1]) and
                                             np.array_equal(lyaps[1][2], lyaps[G][2]))
        res["rk4_speedup"] = res["rk4_e2e_%d" % G] / res["rk4_e2e_1"]
        res["lyapunov_speedup"] = res["lyapunov_e2e_%d" % G] / res["lyapunov_e2e_1"]
    print(json.dumps(res), flush=True)
    if out:
        with open(out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
