#!/usr/bin/env python
"""Golden vectors of the callers either side of the hot path, produced by the UNMODIFIED reference (build container only):

* ``qgs.toolbox.lyapunov._compute_clv_gin_jit`` (lyapunov.py:1174-1288) -- covariant Lyapunov vectors, Ginelli's
  method, with and without the diagonal noise.  Its random start matrices and noise come from numba's generator; that
  generator is seeded from inside a jitted function and the very same draws are taken again, in the same order, and
  stored (``q_draw``, ``a_draw``, ``noise``), so that the device path can be fed identical inputs;
* ``qgs.toolbox.lyapunov._compute_clv_sub_jit`` (lyapunov.py:1292-1329) -- CLVs by subspace intersection, same seeding trick;
* ``qgs.integrators.statistics.TrajectoriesStatistics.compute_stats`` (statistics.py:33-66) on top of the reference's
  own ``RungeKuttaIntegrator`` worker pool.

    python tests/golden/make_golden_extra.py        ->  tests/golden/golden_extra_rp.npz
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("QGS_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "qgs_b200", "compat"))
sys.path.insert(0, REFERENCE)
warnings.filterwarnings("ignore")

from numba import njit  # noqa: E402

from make_golden import closures  # noqa: E402
from qgs.toolbox import lyapunov as ref_lyap  # noqa: E402
from qgs.integrators.integrator import RungeKuttaIntegrator  # noqa: E402
from qgs.integrators.statistics import TrajectoriesStatistics  # noqa: E402


@njit
def _seed(s):
    np.random.seed(s)


@njit
def _randn2(n, m):
    return np.random.randn(n, m)


@njit
def _randn1(n):
    return np.random.randn(n)


@njit
def _random2(n, m):
    return np.random.random((n, m))


def main():
    z = np.load(os.path.join(HERE, "tensor_rp.npz"))
    n = int(z["ndim"])
    f, Df = closures(z["coo"].astype(np.int64), z["val"], z["jcoo"].astype(np.int64), z["jval"], int(z["rank"]))
    rng = np.random.default_rng(4242)
    c = np.array([0., 0.5, 0.5, 1.])
    b = np.array([1. / 6, 1. / 3, 1. / 3, 1. / 6])
    a = np.zeros((4, 4))
    a[1, 0] = 0.5
    a[2, 1] = 0.5
    a[3, 2] = 1.
    out = {}

    # ---- CLVs: one trajectory, so the order of the generator's draws is q-draw, a-draw, then one noise vector per step
    ic = rng.random((1, n)) * 0.1
    pretime = np.concatenate((np.arange(0., 1., 0.1), [1.]))
    time = np.concatenate((np.arange(1., 2.45, 0.1), [2.45]))         # ragged last step
    aftertime = np.concatenate((np.arange(2.45, 3.5, 0.1), [3.5]))
    tw, tew = len(time) - 1, len(time) + len(aftertime) - 2
    out.update(clv_ic=ic, clv_pretime=pretime, clv_time=time, clv_aftertime=aftertime)
    for tag, ws, noise_pert in (("plain", 3, 0.), ("noise", 1, 1e-3)):
        _seed(777)
        q_draw = _randn2(n, n)
        a_draw = _randn2(n, n)
        noise = np.array([_randn1(n) for _ in range(tew)])             # first draw belongs to ti = tew - 1
        _seed(777)
        rt, re, rv = ref_lyap._compute_clv_gin_jit(f, Df, pretime, time, aftertime, 0.1, ic, n, ws, b, c, a, noise_pert)
        out["clv_%s_meta" % tag] = np.array([ws, noise_pert])
        out["clv_%s_q_draw" % tag], out["clv_%s_a_draw" % tag] = q_draw, a_draw
        out["clv_%s_noise" % tag] = noise[::-1].copy()                 # indexed by ti
        out["clv_%s_traj" % tag], out["clv_%s_exp" % tag], out["clv_%s_vec" % tag] = rt, re, rv

    # ---- CLVs by subspace intersection (lyapunov.py:1292-1329): the FLV pass draws its start basis first, then the BLV pass
    _seed(999)
    f_draw = _random2(n, n)
    b_draw = _random2(n, n)
    _seed(999)
    rt, re, rv, bvec, fvec = ref_lyap._compute_clv_sub_jit(f, Df, pretime, time, aftertime, 0.1, ic, 2, b, c, a)
    out["sub_f_draw"], out["sub_b_draw"] = f_draw, b_draw
    out["sub_traj"], out["sub_exp"], out["sub_vec"], out["sub_bvec"], out["sub_fvec"] = rt, re, rv, bvec, fvec

    # ---- TrajectoriesStatistics on the reference's own integrator pool
    sic = rng.random((13, n)) * 0.1
    integ = RungeKuttaIntegrator(num_threads=2)
    integ.set_func(f)
    st = TrajectoriesStatistics()
    st.set_integrator(integ)
    st.set_func_list([lambda x: x, lambda x: x ** 2])
    st.compute_stats(0., 1.25, 0.1, ic=sic, write_steps=4, num=3)
    out["stats_ic"] = sic
    out["stats_mean_func"] = st.get_stats()
    integ.terminate()

    path = os.path.join(HERE, "golden_extra_rp.npz")
    np.savez_compressed(path, **out)
    print("%d arrays, %.1f KB -> %s" % (len(out), os.path.getsize(path) / 1024, path))


if __name__ == "__main__":
    main()
