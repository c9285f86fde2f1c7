#!/usr/bin/env python
"""Generate golden input/output vectors of the hot path by running the UNMODIFIED reference.

Build-container only (needs /root/reference and the tensors from make_tensors.py):

    python tests/golden/make_golden.py [config ...]

What runs is the reference's own numba code: ``qgs.functions.sparse_mul.sparse_mul2/3/4/5``,
``qgs.integrators.integrate._integrate_runge_kutta_jit`` / ``_integrate_runge_kutta_tgls_jit`` and
their Python wrappers, ``qgs.toolbox.lyapunov._compute_backward_lyap_jit`` /
``_compute_forward_lyap_jit``.  The ``f`` / ``Df`` closures are rebuilt from the stored tensors with
the four lines of ``qgs/functions/tendencies.py:98-121`` (so that the 10-minute T4 tensor is not
rebuilt); for the analytic configurations the result is asserted bit-identical to the closures
returned by the reference's ``create_tendencies``.

The Benettin start ``q = qr(random((n_dim, n_vec)))[0]`` (``lyapunov.py:592-593``) comes from
numba's per-process generator; it is made reproducible by seeding that generator from inside a
jitted function, and the same draw is stored in the fixture as ``q0`` / ``r0``.

Outputs: ``tests/golden/golden_<config>.npz``.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("QGS_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qgs_b200", "compat"))
sys.path.insert(0, REFERENCE)
warnings.filterwarnings("ignore")

from numba import njit  # noqa: E402

from qgs.functions.sparse_mul import sparse_mul2, sparse_mul3, sparse_mul4, sparse_mul5  # noqa: E402
from qgs.integrators import integrate as ref_integrate  # noqa: E402
from qgs.toolbox import lyapunov as ref_lyap  # noqa: E402

T4_NOTEBOOK_IC = np.array([
    4.64467907e-02, -2.25054437e-02, -3.51762104e-02, 4.65057475e-03,
    -5.08842464e-03, 1.92789965e-02, 3.47849539e-03, 1.86038254e-02,
    -2.99742925e-03, 8.85862363e-03, 1.53577181e+00, 4.29584574e-02,
    -1.88143295e-02, -1.15114614e-02, 1.24843108e-03, -3.61467108e-03,
    2.81459020e-03, 5.59630451e-03, 3.56804517e-03, -1.52654934e-03,
    3.09046486e-03, 1.18185262e-06, 6.76756261e-04, 4.21981618e-06,
    1.06290914e-05, -1.64056994e-05, 3.40932989e-05, 1.90943692e-05,
    1.08189600e-05, 3.16661484e+00, -3.06749318e-04, 1.82036792e-01,
    -1.33924039e-04, 9.77967977e-03, -5.58853689e-03, 2.33218135e-02,
    3.45971396e-05, -5.88894363e-05
])  # notebooks/maooam_T4.ipynb, cell "ic = np.array([...])"


def closures(coo, val, jcoo, jval, rank):
    """tendencies.py:98-121 verbatim in structure, on stored arrays."""
    if rank == 5:
        @njit
        def f(t, x):
            xx = np.concatenate((np.full((1,), 1.), x))
            xr = sparse_mul5(coo, val, xx, xx, xx, xx)
            return xr[1:]

        @njit
        def Df(t, x):
            xx = np.concatenate((np.full((1,), 1.), x))
            mul_jac = sparse_mul4(jcoo, jval, xx, xx, xx)
            return mul_jac[1:, 1:]
    else:
        @njit
        def f(t, x):
            xx = np.concatenate((np.full((1,), 1.), x))
            xr = sparse_mul3(coo, val, xx, xx)
            return xr[1:]

        @njit
        def Df(t, x):
            xx = np.concatenate((np.full((1,), 1.), x))
            mul_jac = sparse_mul2(jcoo, jval, xx)
            return mul_jac[1:, 1:]
    return f, Df


@njit
def _seed_numba(s):
    np.random.seed(s)


@njit
def _draw_qr(n_dim, n_vec):
    qr = np.linalg.qr(np.random.random((n_dim, n_vec)))
    return qr[0], qr[1]


def states(name, n, count, rng):
    if name in ("T4", "dynT"):
        return T4_NOTEBOOK_IC[None, :] + 1e-3 * rng.standard_normal((count, n))
    if name in ("rp", "tlad"):
        return rng.random((count, n)) * 0.1
    return rng.random((count, n)) * 0.01


SIZES = {   # config: (n_states, n_members, n_steps, tgls_members, tgls_steps, lyap?)
    "rp": (8, 3, 200, 2, 40, True),
    "tlad": (4, 2, 100, 2, 20, False),
    "maooam36": (8, 3, 200, 2, 30, True),
    "aotensor_ref": (4, 2, 50, 1, 10, False),
    "dynT": (6, 2, 100, 2, 20, False),
    "T4": (4, 2, 40, 1, 6, False),
    "atm6x6": (3, 2, 20, 1, 3, False),
}


def make(name):
    z = np.load(os.path.join(HERE, "tensor_%s.npz" % name))
    n, rank = int(z["ndim"]), int(z["rank"])
    coo, val = z["coo"].astype(np.int64), z["val"]
    jcoo, jval = z["jcoo"].astype(np.int64), z["jval"]
    f, Df = closures(coo, val, jcoo, jval, rank)
    rng = np.random.default_rng(21217 + sum(map(ord, name)))
    n_states, n_mem, n_steps, tg_mem, tg_steps, do_lyap = SIZES[name]
    out = {}

    # ---- a5: f, Df on random states --------------------------------------------------------------
    X = states(name, n, n_states, rng)
    out["X"] = X
    out["fX"] = np.array([f(0., x) for x in X])
    out["DfX"] = np.array([Df(0., x) for x in X])

    if name in ("rp", "maooam36"):
        import make_tensors
        from qgs.functions.tendencies import create_tendencies
        fr, Dfr = create_tendencies(make_tensors.CONFIGS[name]())
        assert all(np.array_equal(fr(0., x), f(0., x)) for x in X), "closure != create_tendencies f"
        assert all(np.array_equal(Dfr(0., x), Df(0., x)) for x in X), "closure != create_tendencies Df"

    # ---- a1-a4: the raw contractions with *different* vectors ------------------------------------
    va, vb, vc, vd = (rng.standard_normal(n + 1) for _ in range(4))
    out["vecs"] = np.array([va, vb, vc, vd])
    if rank == 3:
        out["mul_vec"] = sparse_mul3(coo, val, va, vb)
        out["mul_mat"] = sparse_mul2(jcoo, jval, va)
    else:
        out["mul_vec"] = sparse_mul5(coo, val, va, vb, vc, vd)
        out["mul_mat"] = sparse_mul4(jcoo, jval, va, vb, vc)

    # ---- a7: _integrate_runge_kutta_jit -----------------------------------------------------------
    c = np.array([0., 0.5, 0.5, 1.])
    b = np.array([1. / 6, 1. / 3, 1. / 3, 1. / 6])
    a = np.zeros((4, 4))
    a[1, 0] = 0.5
    a[2, 1] = 0.5
    a[3, 2] = 1.
    dt = 0.1
    ic = states(name, n, n_mem, rng)
    out["ic"] = ic
    time = np.concatenate((np.arange(0., n_steps * dt, dt), np.full((1,), n_steps * dt)))
    out["rk_time"] = time
    for tag, (direction, ws) in {"fwd_ws10": (1, 10), "fwd_ws0": (1, 0), "fwd_ws7": (1, 7),
                                 "bwd_ws3": (-1, 3), "fwd_ws1": (1, 1)}.items():
        out["rk_" + tag] = ref_integrate._integrate_runge_kutta_jit(f, time, ic, direction, ws, b, c, a)
    # a generic tableau (Kutta 3/8 rule) and a 2-stage method (Heun) on a ragged time vector
    c38 = np.array([0., 1. / 3, 2. / 3, 1.])
    b38 = np.array([1. / 8, 3. / 8, 3. / 8, 1. / 8])
    a38 = np.array([[0., 0., 0., 0.], [1. / 3, 0., 0., 0.], [-1. / 3, 1., 0., 0.], [1., -1., 1., 0.]])
    tr = np.concatenate((np.arange(0.3, 0.3 + 0.07 * 13.5, 0.07), np.full((1,), 0.3 + 0.07 * 13.5)))
    out["rk38_time"] = tr
    out["rk38"] = ref_integrate._integrate_runge_kutta_jit(f, tr, ic, 1, 4, b38, c38, a38)
    ch, bh, ah = np.array([0., 1.]), np.array([0.5, 0.5]), np.array([[0., 0.], [1., 0.]])
    out["rkheun"] = ref_integrate._integrate_runge_kutta_jit(f, tr, ic, -1, 2, bh, ch, ah)

    # ---- a8: the functional wrapper (time vector rules) ------------------------------------------
    for tag, kw in {"w_fwd_ws4": dict(forward=True, write_steps=4),
                    "w_bwd_ws4": dict(forward=False, write_steps=4),
                    "w_fwd_ws0": dict(forward=True, write_steps=0),
                    "w_bwd_ws5": dict(forward=False, write_steps=5)}.items():
        tt, tj = ref_integrate.integrate_runge_kutta(f, 0., 1.35, 0.1, ic=ic, **kw)
        out[tag + "_t"] = np.asarray(tt)
        out[tag + "_x"] = tj
    tt, tj = ref_integrate.integrate_runge_kutta(f, 0., 1., 0.1, ic=ic[0], write_steps=2)
    out["w_single_t"], out["w_single_x"] = tt, tj

    # ---- a9: _integrate_runge_kutta_tgls_jit ------------------------------------------------------
    tic = ic[:tg_mem]
    ttime = np.concatenate((np.arange(0., tg_steps * dt, dt), np.full((1,), tg_steps * dt)))
    out["tg_time"] = ttime
    Id = np.repeat(np.eye(n)[None], tg_mem, axis=0)
    out["tg_id_traj"], out["tg_id_fm"] = ref_integrate._integrate_runge_kutta_tgls_jit(
        f, Df, ttime, tic, Id, 1, 5, b, c, a, False, 1., ref_integrate._zeros_func)
    tg3 = rng.standard_normal((tg_mem, n, 3))
    out["tg3_ic"] = tg3
    out["tg3_traj"], out["tg3_fm"] = ref_integrate._integrate_runge_kutta_tgls_jit(
        f, Df, ttime, tic, tg3, 1, 0, b, c, a, False, 1., ref_integrate._zeros_func)
    out["tg3_adj_traj"], out["tg3_adj_fm"] = ref_integrate._integrate_runge_kutta_tgls_jit(
        f, Df, ttime, tic, tg3, -1, 4, b, c, a, True, -1., ref_integrate._zeros_func)
    # the wrapper's tg_ic shape heuristics (integrate.py:479-500)
    if name in ("rp", "maooam36"):
        v1 = rng.standard_normal(n)
        for tag, tg in {"wt_none": None, "wt_1d": v1, "wt_2d_ens": rng.standard_normal((4, n)),
                        "wt_2d_per": rng.standard_normal((tg_mem, n)),
                        "wt_3d": rng.standard_normal((tg_mem, 4, n))}.items():
            if tg is not None:
                out[tag + "_tg"] = tg
            r = ref_integrate.integrate_runge_kutta_tgls(f, Df, 0., 0.5, 0.1, ic=tic, tg_ic=tg, write_steps=2)
            out[tag + "_t"], out[tag + "_x"], out[tag + "_fm"] = r

    # ---- a13: Benettin BLV / FLV ---------------------------------------------------------------------
    if do_lyap:
        lic = ic[:2]
        for tag, (nv, mdt, ws) in {"full": (n, 0.1, 3), "sub": (5, 0.05, 2)}.items():
            pretime = np.concatenate((np.arange(0., 1., dt), np.full((1,), 1.)))
            ltime = np.concatenate((np.arange(1., 2.5, dt), np.full((1,), 2.5)))
            _seed_numba(1234)
            q0 = np.zeros((2, n, nv))
            r0 = np.zeros((2, nv, nv))
            for i in range(2):
                q0[i], r0[i] = _draw_qr(n, nv)
            _seed_numba(1234)
            rt, re, rv = ref_lyap._compute_backward_lyap_jit(f, Df, pretime, ltime, mdt, lic, nv, ws,
                                                            False, 1., b, c, a)
            out["blv_%s_q0" % tag], out["blv_%s_r0" % tag] = q0, r0
            out["blv_%s_traj" % tag], out["blv_%s_exp" % tag], out["blv_%s_vec" % tag] = rt, re, rv
            _seed_numba(1234)
            rt, re, rv = ref_lyap._compute_forward_lyap_jit(f, Df, pretime, ltime, mdt, lic, nv, ws,
                                                           False, 1., b, c, a)
            out["flv_%s_traj" % tag], out["flv_%s_exp" % tag], out["flv_%s_vec" % tag] = rt, re, rv
            out["lyap_%s_meta" % tag] = np.array([nv, mdt, ws])
        out["lyap_pretime"], out["lyap_time"] = pretime, ltime

    path = os.path.join(HERE, "golden_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-12s %3d arrays, %.1f KB -> %s" % (name, len(out), os.path.getsize(path) / 1024, path), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    for nm in (sys.argv[1:] or list(SIZES)):
        if os.path.exists(os.path.join(HERE, "tensor_%s.npz" % nm)):
            make(nm)
        else:
            print("skip %s (no tensor yet)" % nm)
