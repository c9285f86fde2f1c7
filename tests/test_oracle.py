"""CPU: pin the oracle (oracle/qgs_oracle.c) against the reference.

Two kinds of pins:
* the reference's own golden files for the *input* of the path (model_test/test_aotensor*.ref,
  converted to tests/golden/refpin_*.npz by tests/golden/make_tensors.py --pins), compared with the
  reference test's own rule (6 printed digits, |diff| < 5 eps on the printed value,
  model_test/test_aotensor_6x6.py:21);
* outputs of the unmodified reference numba code (tests/golden/golden_*.npz, make_golden.py).
"""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["rp", "tlad", "maooam36", "aotensor_ref", "dynT", "T4", "atm6x6"]


def load(name):
    T = oracle.Tensor.from_npz(os.path.join(GOLDEN, "tensor_%s.npz" % name))
    g = np.load(os.path.join(GOLDEN, "golden_%s.npz" % name))
    return T, g


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("tensor,pin,jac", [("aotensor_ref", "refpin_aotensor", False),
                                            ("aotensor_ref", "refpin_aotensor_jacobian", True),
                                            ("atm6x6", "refpin_aotensor_6x6", False)])
def test_tensor_fixture_matches_reference_ref_files(tensor, pin, jac):
    z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % tensor))
    p = np.load(os.path.join(GOLDEN, pin + ".npz"))
    coo, val = (z["jcoo"], z["jval"]) if jac else (z["coo"], z["val"])
    mine = {tuple(int(i) for i in c): v for c, v in zip(coo, val)}
    eps = 5 * np.finfo(np.float64).eps
    assert len(p["val"]) >= 0.999 * len(val)
    for c, v in zip(p["coo"], p["val"]):
        printed = float("% .5E" % mine[tuple(int(i) for i in c)])
        assert abs(printed - v) < eps
    # entries absent from the .ref are those the reference test drops (|value| < 5 eps)
    missing = set(mine) - {tuple(int(i) for i in c) for c in p["coo"]}
    assert all(abs(mine[k]) < eps for k in missing)


@pytest.mark.parametrize("name", CONFIGS)
def test_tendencies_and_jacobian(name):
    T, g = load(name)
    assert rel(oracle.f(T, g["X"]), g["fX"]) < 1e-14
    assert rel(oracle.Df(T, g["X"]), g["DfX"]) < 1e-14


@pytest.mark.parametrize("name", CONFIGS)
def test_sparse_mul_distinct_vectors(name):
    T, g = load(name)
    va, vb, vc, vd = g["vecs"]
    if T.rank == 3:
        v = oracle.sparse_mul3(T.coo, T.val, va, vb)
        m = oracle.sparse_mul2(T.jcoo, T.jval, va)
    else:
        v = oracle.sparse_mul5(T.coo, T.val, va, vb, vc, vd)
        m = oracle.sparse_mul4(T.jcoo, T.jval, va, vb, vc)
    assert v[0] == 1.0
    assert rel(v, g["mul_vec"]) < 1e-14
    assert rel(m, g["mul_mat"]) < 1e-14


@pytest.mark.parametrize("name", CONFIGS)
def test_rk_trajectories(name):
    T, g = load(name)
    b, c, a = oracle.rk4_tableau()
    for tag, (direction, ws) in {"fwd_ws10": (1, 10), "fwd_ws0": (1, 0), "fwd_ws7": (1, 7),
                                 "bwd_ws3": (-1, 3), "fwd_ws1": (1, 1)}.items():
        ref = g["rk_" + tag]
        got = oracle.integrate_runge_kutta_jit(T, g["rk_time"], g["ic"], direction, ws, b, c, a)
        assert got.shape == ref.shape
        assert rel(got, ref) < 1e-12, tag
    c38 = np.array([0., 1. / 3, 2. / 3, 1.])
    b38 = np.array([1. / 8, 3. / 8, 3. / 8, 1. / 8])
    a38 = np.array([[0., 0., 0., 0.], [1. / 3, 0., 0., 0.], [-1. / 3, 1., 0., 0.], [1., -1., 1., 0.]])
    got = oracle.integrate_runge_kutta_jit(T, g["rk38_time"], g["ic"], 1, 4, b38, c38, a38)
    assert got.shape == g["rk38"].shape and rel(got, g["rk38"]) < 1e-12
    ch, bh, ah = np.array([0., 1.]), np.array([0.5, 0.5]), np.array([[0., 0.], [1., 0.]])
    got = oracle.integrate_runge_kutta_jit(T, g["rk38_time"], g["ic"], -1, 2, bh, ch, ah)
    assert got.shape == g["rkheun"].shape and rel(got, g["rkheun"]) < 1e-12


@pytest.mark.parametrize("name", CONFIGS)
def test_tgls(name):
    T, g = load(name)
    b, c, a = oracle.rk4_tableau()
    n = T.ndim
    tic = g["ic"][:g["tg3_ic"].shape[0]]
    Id = np.repeat(np.eye(n)[None], len(tic), axis=0)
    traj, fm = oracle.integrate_runge_kutta_tgls_jit(T, g["tg_time"], tic, Id, 1, 5, b, c, a, False, 1.)
    assert traj.shape == g["tg_id_traj"].shape and fm.shape == g["tg_id_fm"].shape
    assert rel(traj, g["tg_id_traj"]) < 1e-12 and rel(fm, g["tg_id_fm"]) < 1e-11
    traj, fm = oracle.integrate_runge_kutta_tgls_jit(T, g["tg_time"], tic, g["tg3_ic"], 1, 0, b, c, a, False, 1.)
    assert rel(traj, g["tg3_traj"]) < 1e-12 and rel(fm, g["tg3_fm"]) < 1e-11
    traj, fm = oracle.integrate_runge_kutta_tgls_jit(T, g["tg_time"], tic, g["tg3_ic"], -1, 4, b, c, a, True, -1.)
    assert rel(traj, g["tg3_adj_traj"]) < 1e-12 and rel(fm, g["tg3_adj_fm"]) < 1e-11


def test_qr_matches_lapack_convention():
    rng = np.random.default_rng(5)
    for n, m in ((36, 36), (36, 5), (20, 20), (7, 3)):
        A = rng.standard_normal((n, m))
        Q, R = oracle.qr(A)
        Qn, Rn = np.linalg.qr(A)
        assert np.allclose(Q, Qn, atol=1e-12) and np.allclose(R, Rn, atol=1e-12)


@pytest.mark.parametrize("name", ["rp", "maooam36"])
@pytest.mark.parametrize("tag", ["full", "sub"])
def test_benettin_blv_flv(name, tag):
    T, g = load(name)
    b, c, a = oracle.rk4_tableau()
    nv, mdt, ws = g["lyap_%s_meta" % tag]
    nv, ws = int(nv), int(ws)
    lic = g["ic"][:2]
    q0, r0 = g["blv_%s_q0" % tag], g["blv_%s_r0" % tag]
    rt, re, rv = oracle.compute_backward_lyap(T, g["lyap_pretime"], g["lyap_time"], mdt, lic, nv, ws,
                                              False, 1., b, c, a, q0, r0)
    assert rt.shape == g["blv_%s_traj" % tag].shape
    assert rel(rt, g["blv_%s_traj" % tag]) < 1e-12
    assert rel(re, g["blv_%s_exp" % tag]) < 1e-8
    assert rel(rv, g["blv_%s_vec" % tag]) < 1e-8
    rt, re, rv = oracle.compute_forward_lyap(T, g["lyap_pretime"], g["lyap_time"], mdt, lic, nv, ws,
                                             False, 1., b, c, a, q0, r0)
    assert rt.shape == g["flv_%s_traj" % tag].shape
    assert rel(rt, g["flv_%s_traj" % tag]) < 1e-12
    assert rel(re, g["flv_%s_exp" % tag]) < 1e-8
    assert rel(rv, g["flv_%s_vec" % tag]) < 1e-8


def test_n_records_rule():
    # integrate.py:190-196
    for L in range(1, 40):
        time = np.arange(L, dtype=float)
        for ws in range(0, 9):
            if ws == 0:
                exp = 1
            else:
                tot = time[::ws]
                exp = len(tot) + (1 if tot[-1] != time[-1] else 0)
            assert oracle.n_records(L, ws) == exp


def test_householder_q_does_not_depend_on_the_signs_of_the_columns():
    """The property the device's Cholesky-QR steps rest on: np.linalg.qr -- LAPACK's Householder
    factorisation, lyapunov.py:602-604 -- returns the same Q for A and for A D, D = diag(+-1), and R with its columns
    flipped.  So an unobserved Benettin step may hand on Q D instead of Q: the next Householder step gives the same Q and
    the same |diag R|."""
    rng = np.random.default_rng(12)
    for n, m in ((36, 36), (36, 10), (20, 20)):
        a = rng.standard_normal((n, m))
        d = np.where(rng.random(m) < 0.5, -1., 1.)
        q0, r0 = np.linalg.qr(a)
        q1, r1 = np.linalg.qr(a * d)
        assert np.abs(q0 - q1).max() < 1e-13
        assert np.abs(r0 * d - r1).max() < 1e-13
        # and a Cholesky QR spans the same nested subspaces: Q_chol = Q D' with D' = sign(diag R)
        rc = np.linalg.cholesky(a.T @ a).T
        qc = a @ np.linalg.inv(rc)
        assert np.abs(qc - q0 * np.sign(np.diag(r0))).max() < 1e-11
        assert np.abs(np.diag(rc) - np.abs(np.diag(r0))).max() < 1e-12
