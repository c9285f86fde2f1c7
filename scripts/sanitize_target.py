"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgs_b200 import _lib  # noqa: E402
from qgs_b200.ensemble import DeviceEnsemble  # noqa: E402
from qgs_b200.functions.tendencies import tendencies_from_tensor  # noqa: E402
from qgs_b200.integrators.integrator import RungeKuttaIntegrator, RungeKuttaTglsIntegrator  # noqa: E402
from qgs_b200.toolbox.lyapunov import CovariantLyapunovsEstimator, LyapunovsEstimator  # noqa: E402

_lib.init(0)
which = sys.argv[1:] or ["rk", "large", "tangent", "clv", "stats", "round2"]


def load(name):
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
    return tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])


rng = np.random.default_rng(0)
if "rk" in which:
    f, Df = load("maooam36")
    integ = RungeKuttaIntegrator()
    integ.set_func(f)
    for rows_max, N in (("2048", 3), ("0", 200)):            # rows kernel; specialised kernel
        os.environ["QGSB_RK_ROWS_MAX"] = rows_max
        integ.integrate(0., 0.5, 0.1, ic=rng.random((N, 36)) * 0.01, write_steps=2)
    f.tensor.use_specialised(False)
    integ.integrate(0., 0.3, 0.1, ic=rng.random((130, 36)) * 0.01, write_steps=0)      # generic G1
    os.environ.pop("QGSB_RK_ROWS_MAX")
    print("rk ok", flush=True)
if "large" in which:
    f, Df = load("atm6x6")
    integ = RungeKuttaIntegrator()
    integ.set_func(f)
    integ.integrate(0., 0.2, 0.1, ic=rng.random((2, 228)) * 0.01, write_steps=1)       # wide kernel
    os.environ["QGSB_RK_ROWS_MAX"] = "0"
    integ.integrate(0., 0.2, 0.1, ic=rng.random((100, 228)) * 0.01, write_steps=1)     # G3
    os.environ["QGSB_RK_LARGE"] = "g2"
    integ.integrate(0., 0.1, 0.1, ic=rng.random((5, 228)) * 0.01, write_steps=0)       # G2
    os.environ.pop("QGSB_RK_ROWS_MAX")
    os.environ.pop("QGSB_RK_LARGE")
    print("large ok", flush=True)
if "tangent" in which:
    f, Df = load("maooam36")
    tg = RungeKuttaTglsIntegrator()
    tg.set_func(f, Df)
    ic = rng.random((9, 36)) * 0.01
    for kern in ("pack", "pack_dense", "generic"):
        os.environ["QGSB_TGLS_KERNEL"] = kern
        tg.integrate(0., 0.3, 0.1, ic=ic, write_steps=1)
        tg.integrate(0., 0.2, 0.1, ic=ic, write_steps=0, adjoint=True, inverse=True)
        est = LyapunovsEstimator()
        est.set_func(f, Df)
        est.compute_lyapunovs(0., 0.2, 0.5, 0.1, 0.05, ic=ic, write_steps=1, n_vec=36 if kern != "generic" else 7)
        est.compute_lyapunovs(0., 0.2, 0.5, 0.1, 0.1, ic=ic, write_steps=2, n_vec=20, forward=True)
    os.environ.pop("QGSB_TGLS_KERNEL")
    print("tangent ok", flush=True)
if "clv" in which:
    f, Df = load("rp")
    ic = rng.random((3, 20)) * 0.1
    for method in (0, 1):
        est = CovariantLyapunovsEstimator(method=method)
        est.set_func(f, Df)
        est.compute_clvs(0., 0.3, 0.8, 1.2, 0.1, 0.1, ic=ic, write_steps=2, method=method)
    print("clv ok", flush=True)
if "stats" in which:
    f, Df = load("rp")
    ens = DeviceEnsemble(f, rng.random((300, 20)) * 0.1)
    os.environ["QGSB_STREAM_RECORDS"] = "2"
    ens.integrate_trajectories(0., 0.7, 0.1, write_steps=1)
    ens.integrate_moments(0.7, 1.2, 0.1, write_steps=2)
    ens.moments()
    f(0., rng.random((600, 20)))
    Df(0., rng.random((4, 20)))
    print("stats ok", flush=True)
if "round2" in which:
    # round 2: general-tableau kernel (half tiles), tail-wave launch, device-drawn start bases (step -1 of the Benettin
    # kernels, packed and generic), member batches, initialize() on resident batches
    f, Df = load("maooam36")
    integ = RungeKuttaIntegrator(b=np.array([1., 3., 3., 1.]) / 8., c=np.array([0., 1. / 3, 2. / 3, 1.]),
                                 a=np.array([[0., 0., 0., 0.], [1. / 3, 0., 0., 0.], [-1. / 3, 1., 0., 0.],
                                             [1., -1., 1., 0.]]))
    integ.set_func(f)
    os.environ["QGSB_RK_ROWS_MAX"] = "0"
    integ.integrate(0., 0.3, 0.1, ic=rng.random((200, 36)) * 0.01, write_steps=2)       # rk_general_kernel
    integ = RungeKuttaIntegrator()
    integ.set_func(f)
    n_tail = (_lib.device_info()["sm_count"] * 2 + 5) * 128                             # one full wave + 5 tail tiles
    integ.integrate(0., 0.2, 0.1, ic=rng.random((n_tail, 36)) * 0.01, write_steps=0)
    os.environ.pop("QGSB_RK_ROWS_MAX")
    integ.device_batch = 50
    integ.initialize(0.3, 0.1, reconvergence_time=0.2, number_of_trajectories=120, ic=rng.random((120, 36)) * 0.01,
                     reconverge=True)
    ic = rng.random((9, 36)) * 0.01
    os.environ["QGSB_TANGENT_BUDGET_MB"] = "1"
    for kern in ("pack", "generic"):
        os.environ["QGSB_TGLS_KERNEL"] = kern
        est = LyapunovsEstimator()
        est.set_func(f, Df)
        est.compute_lyapunovs(0., 0.2, 0.5, 0.1, 0.1, ic=ic, write_steps=1, n_vec=36 if kern == "pack" else 5)
        est.compute_lyapunovs(0., 0.2, 0.5, 0.1, 0.1, ic=ic, write_steps=2, n_vec=12, vectors=False)
    os.environ.pop("QGSB_TGLS_KERNEL")
    os.environ.pop("QGSB_TANGENT_BUDGET_MB")
    print("round2 ok", flush=True)

if "cholqr" in which:
    # Cholesky QR of the packed Benettin kernel (pack::chol_factor / chol_solve): every column capacity, with and
    # without vector records (write_steps = 3: two Cholesky steps, one Householder step), a basis with a repeated column
    # (pivot refused, per-member Householder fallback), the tangent-linear kernel with the stage sum in registers
    from qgs_b200.toolbox.lyapunov import benettin
    from qgs_b200.integrators.integrate import rk4_tableau
    b_, c_, a_ = rk4_tableau()
    pre = np.concatenate((np.arange(0., 0.3, 0.1), [0.3]))
    tim = np.concatenate((np.arange(0.3, 1.0, 0.1), [1.0]))
    for name, vecs in (("maooam36", (36, 20, 10)), ("rp", (20, 12)), ("dynT", (38, 30))):
        f, Df = load(name)
        n = f.ndim
        ic = rng.random((9, n)) * 0.01
        for m in vecs:
            q0 = np.stack([np.linalg.qr(rng.random((n, m)))[0] for _ in range(9)])
            for vectors in (True, False):
                benettin(f, Df, ic, 0, m, q0, None, pre, tim, 0.1, 3, False, 1., b_, c_, a_, want_vectors=vectors)
        q0 = np.stack([np.linalg.qr(rng.random((n, n)))[0] for _ in range(9)])
        q0[4, :, 3] = q0[4, :, 2]
        benettin(f, Df, ic, 0, n, q0, None, pre, tim, 0.1, 3, False, 1., b_, c_, a_, want_vectors=False)
        tg = RungeKuttaTglsIntegrator()
        tg.set_func(f, Df)
        tg.integrate(0., 0.3, 0.1, ic=ic, write_steps=1)
    print("cholqr ok", flush=True)
