"""GPU parity tests: the CUDA path (through the C ABI, via the Python mirror of the reference API)
against (a) the golden vectors produced by the unmodified reference (tests/golden/golden_*.npz) and
(b) the CPU oracle on seeded inputs.

Tolerances (float64, stated by BASELINE.json's north_star / SURVEY.md section 8c):
  * tendencies f and Jacobian Df: 1e-12 relative to max|f| (resp. max|Df|) of the state;
  * trajectories: 1e-10 * max|y| over the (short) golden horizons;
  * tangent-linear fundamental matrices: 1e-10 relative;
  * Benettin exponents / vectors over the golden windows: 1e-8 relative (same start basis).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["rp", "tlad", "maooam36", "aotensor_ref", "dynT", "T4", "atm6x6"]
_cache = {}


def setup(name, specialise=True):
    key = (name, specialise)
    if key not in _cache:
        from qgs_b200.functions.tendencies import tendencies_from_tensor
        z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        g = np.load(os.path.join(GOLDEN, "golden_%s.npz" % name))
        f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"], specialise=specialise)
        if not specialise:
            f.tensor.use_specialised(False)
        _cache[key] = (f, Df, g, z)
    return _cache[key]


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def rk4():
    from qgs_b200.integrators.integrate import rk4_tableau
    return rk4_tableau()


# ---- a1-a4: raw contractions -------------------------------------------------------------------------
@pytest.mark.parametrize("name", CONFIGS)
def test_sparse_mul_raw(name):
    from qgs_b200.functions import sparse_mul as sm
    _, _, g, z = setup(name)
    va, vb, vc, vd = g["vecs"]
    if int(z["rank"]) == 3:
        v = sm.sparse_mul3(z["coo"], z["val"], va, vb)
        m = sm.sparse_mul2(z["jcoo"], z["jval"], va)
    else:
        v = sm.sparse_mul5(z["coo"], z["val"], va, vb, vc, vd)
        m = sm.sparse_mul4(z["jcoo"], z["jval"], va, vb, vc)
    assert v[0] == 1.0
    assert rel(v, g["mul_vec"]) < 1e-12
    assert rel(m, g["mul_mat"]) < 1e-12


def test_raw_contractions_reuse_the_device_tensor_by_content():
    """sparse_mul2/3 take the tensor as host arrays on every call; the library keeps the prepared device copy in a cache
    keyed by the CONTENT of (coo, val).  Repeated calls, a tensor rewritten in place, arrays that only share their
    content, and more distinct tensors than the cache holds must all give what a fresh call gives."""
    from qgs_b200.functions import sparse_mul as sm
    rng = np.random.default_rng(6)
    n1 = 13

    def rand_tensor(nnz):
        coo = np.stack([rng.integers(0, n1, nnz) for _ in range(3)], axis=1)
        return coo, rng.standard_normal(nnz)

    def expect3(coo, val, a, b):
        out = np.zeros(n1)
        np.add.at(out, coo[:, 0], val * a[coo[:, 1]] * b[coo[:, 2]])
        out[0] = 1.
        return out

    def expect2(coo, val, a):
        out = np.zeros((n1, n1))
        np.add.at(out, (coo[:, 0], coo[:, 1]), val * a[coo[:, 2]])
        return out

    a, b = rng.standard_normal(n1), rng.standard_normal(n1)
    coo, val = rand_tensor(60)
    for _ in range(3):                                               # a hit after the first call
        assert np.allclose(sm.sparse_mul3(coo, val, a, b), expect3(coo, val, a, b), rtol=1e-13, atol=1e-13)
        assert np.allclose(sm.sparse_mul2(coo, val, a), expect2(coo, val, a), rtol=1e-13, atol=1e-13)
    val[7] = 42.                                                     # the same arrays, another content
    assert np.allclose(sm.sparse_mul3(coo, val, a, b), expect3(coo, val, a, b), rtol=1e-13, atol=1e-13)
    coo[3] = (5, 2, 9)
    assert np.allclose(sm.sparse_mul2(coo, val, a), expect2(coo, val, a), rtol=1e-13, atol=1e-13)
    assert np.allclose(sm.sparse_mul3(coo.copy(), val.copy(), a, b), expect3(coo, val, a, b), rtol=1e-13, atol=1e-13)
    tensors = [rand_tensor(20 + q) for q in range(40)]               # more than the cache keeps, twice round
    for _ in range(2):
        for c, v in tensors:
            assert np.allclose(sm.sparse_mul3(c, v, a, b), expect3(c, v, a, b), rtol=1e-13, atol=1e-13)


def test_sparse_mul_edge_cases():
    from qgs_b200.functions import sparse_mul as sm
    # empty tensor: res = 0 except res[0] = 1
    v = sm.sparse_mul3(np.zeros((0, 3), dtype=int), np.zeros(0), np.ones(4), np.ones(4))
    assert np.array_equal(v, [1., 0., 0., 0.])
    assert np.array_equal(sm.sparse_mul2(np.zeros((0, 3), dtype=int), np.zeros(0), np.ones(3)), np.zeros((3, 3)))
    # unsorted input with duplicates and row-0 entries
    coo = np.array([[2, 1, 1], [0, 1, 2], [1, 0, 0], [2, 1, 1], [1, 2, 2]])
    val = np.array([1.5, 7., -2., 0.5, 3.])
    a, b = np.array([1., 2., 3.]), np.array([1., -1., 4.])
    exp = np.zeros(3)
    for (i, j, k), w in zip(coo, val):
        exp[i] += a[j] * b[k] * w
    exp[0] = 1.
    assert np.allclose(sm.sparse_mul3(coo, val, a, b), exp, rtol=1e-15)
    with pytest.raises(RuntimeError):
        sm.sparse_mul3(np.array([[1, 5, 0]]), np.array([1.]), a, b)   # index out of range


# ---- a5: f / Df ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CONFIGS)
def test_tendencies_and_jacobian_vs_golden(name):
    f, Df, g, _ = setup(name)
    X = g["X"]
    assert rel(f(0., X), g["fX"]) < 1e-12
    assert rel(Df(0., X), g["DfX"]) < 1e-12
    assert rel(f(0., X[0]), g["fX"][0]) < 1e-12          # single state, like the reference closure
    assert Df(0., X[0]).shape == (X.shape[1], X.shape[1])
    with pytest.raises(ValueError):
        f(0., np.zeros(X.shape[1] + 1))


@pytest.mark.parametrize("name", ["maooam36", "T4"])
def test_tendencies_vs_oracle_random_states(name):
    import oracle
    f, Df, g, _ = setup(name)
    T = oracle.Tensor.from_npz(os.path.join(GOLDEN, "tensor_%s.npz" % name))
    rng = np.random.default_rng(7)
    X = g["X"][0][None, :] + 1e-2 * rng.standard_normal((513, T.ndim))
    ref = oracle.f(T, X)
    got = f(0., X)
    per_state = np.max(np.abs(got - ref), axis=1) / np.max(np.abs(ref), axis=1)
    assert per_state.max() < 1e-12


# ---- a7/a8: Runge-Kutta trajectories ------------------------------------------------------------------
RK_CASES = {"fwd_ws10": (1, 10), "fwd_ws0": (1, 0), "fwd_ws7": (1, 7), "bwd_ws3": (-1, 3), "fwd_ws1": (1, 1)}


@pytest.mark.parametrize("specialise", [True, False])
@pytest.mark.parametrize("name", CONFIGS)
def test_rk_trajectories_vs_golden(name, specialise):
    from qgs_b200.integrators.integrate import _integrate_runge_kutta_jit
    f, _, g, _ = setup(name, specialise)
    if specialise and f.tensor.kernel_kind != 2:
        pytest.skip("no specialised kernel for this tensor")
    b, c, a = rk4()
    # few members take the block-per-member "rows" kernel by default; QGSB_RK_ROWS_MAX=0 sends the same members
    # through the throughput kernels (tensor-specialised / thread per member / large basis), so both are checked
    for rows_max in ("0", "2048"):
        os.environ["QGSB_RK_ROWS_MAX"] = rows_max
        try:
            for tag, (direction, ws) in RK_CASES.items():
                got = _integrate_runge_kutta_jit(f, g["rk_time"], g["ic"], direction, ws, b, c, a)
                assert rel(got, g["rk_" + tag]) < 1e-10, (tag, rows_max)
        finally:
            del os.environ["QGSB_RK_ROWS_MAX"]
    os.environ["QGSB_RK_ROWS_MAX"] = "0"        # the generic-tableau cases below: throughput kernels
    # generic tableaux on a ragged time vector: Kutta 3/8 (not a chain) and Heun (2 stages, backward)
    c38 = np.array([0., 1. / 3, 2. / 3, 1.])
    b38 = np.array([1. / 8, 3. / 8, 3. / 8, 1. / 8])
    a38 = np.array([[0., 0., 0., 0.], [1. / 3, 0., 0., 0.], [-1. / 3, 1., 0., 0.], [1., -1., 1., 0.]])
    got = _integrate_runge_kutta_jit(f, g["rk38_time"], g["ic"], 1, 4, b38, c38, a38)
    assert rel(got, g["rk38"]) < 1e-10
    ch, bh, ah = np.array([0., 1.]), np.array([0.5, 0.5]), np.array([[0., 0.], [1., 0.]])
    got = _integrate_runge_kutta_jit(f, g["rk38_time"], g["ic"], -1, 2, bh, ch, ah)
    assert rel(got, g["rkheun"]) < 1e-10
    del os.environ["QGSB_RK_ROWS_MAX"]
    got = _integrate_runge_kutta_jit(f, g["rk38_time"], g["ic"], -1, 2, bh, ch, ah)      # Heun on the rows kernel
    assert rel(got, g["rkheun"]) < 1e-10


def test_specialised_kernels_are_linked_for_canonical_tensors():
    for name in ("rp", "maooam36", "dynT"):
        f, _, _, _ = setup(name)
        assert f.tensor.kernel_kind == 2, name


@pytest.mark.parametrize("name", ["rp", "maooam36"])
def test_functional_wrapper_time_vectors(name):
    from qgs_b200.integrators.integrate import integrate_runge_kutta
    f, _, g, _ = setup(name)
    for tag, kw in {"w_fwd_ws4": dict(forward=True, write_steps=4), "w_bwd_ws4": dict(forward=False, write_steps=4),
                    "w_fwd_ws0": dict(forward=True, write_steps=0),
                    "w_bwd_ws5": dict(forward=False, write_steps=5)}.items():
        tt, tj = integrate_runge_kutta(f, 0., 1.35, 0.1, ic=g["ic"], **kw)
        assert np.array_equal(np.asarray(tt), g[tag + "_t"]), tag
        assert rel(tj, g[tag + "_x"]) < 1e-10, tag
    tt, tj = integrate_runge_kutta(f, 0., 1., 0.1, ic=g["ic"][0], write_steps=2)
    assert np.array_equal(tt, g["w_single_t"]) and rel(tj, g["w_single_x"]) < 1e-10


def test_ragged_ensemble_sizes_vs_oracle():
    """Member counts around the 128-wide tile and a single member."""
    import oracle
    from qgs_b200.integrators.integrate import _integrate_runge_kutta_jit
    f, _, g, _ = setup("maooam36")
    T = oracle.Tensor.from_npz(os.path.join(GOLDEN, "tensor_maooam36.npz"))
    b, c, a = rk4()
    rng = np.random.default_rng(3)
    time = np.concatenate((np.arange(0., 2., 0.1), [2.]))
    for N in (1, 127, 128, 129, 300):
        ic = rng.random((N, 36)) * 0.01
        ref = oracle.integrate_runge_kutta_jit(T, time, ic, 1, 5, b, c, a)
        got = _integrate_runge_kutta_jit(f, time, ic, 1, 5, b, c, a)
        assert rel(got, ref) < 1e-10, N
    # zero steps: time = [t0] only -> one record holding the initial condition
    got = _integrate_runge_kutta_jit(f, np.array([0.]), ic, 1, 1, b, c, a)
    assert got.shape == (300, 36, 1) and np.array_equal(got[:, :, 0], ic)


# ---- a9: tangent linear / adjoint ------------------------------------------------------------------------
@pytest.mark.parametrize("name", CONFIGS)
def test_tgls_vs_golden(name):
    from qgs_b200.integrators.integrate import _integrate_runge_kutta_tgls_jit, _zeros_func
    f, Df, g, _ = setup(name)
    b, c, a = rk4()
    n = f.ndim
    tic = g["ic"][:g["tg3_ic"].shape[0]]
    Id = np.repeat(np.eye(n)[None], len(tic), axis=0)
    traj, fm = _integrate_runge_kutta_tgls_jit(f, Df, g["tg_time"], tic, Id, 1, 5, b, c, a, False, 1., _zeros_func)
    assert rel(traj, g["tg_id_traj"]) < 1e-10 and rel(fm, g["tg_id_fm"]) < 1e-10
    traj, fm = _integrate_runge_kutta_tgls_jit(f, Df, g["tg_time"], tic, g["tg3_ic"], 1, 0, b, c, a, False, 1., None)
    assert rel(traj, g["tg3_traj"]) < 1e-10 and rel(fm, g["tg3_fm"]) < 1e-10
    traj, fm = _integrate_runge_kutta_tgls_jit(f, Df, g["tg_time"], tic, g["tg3_ic"], -1, 4, b, c, a, True, -1., None)
    assert rel(traj, g["tg3_adj_traj"]) < 1e-10 and rel(fm, g["tg3_adj_fm"]) < 1e-10


@pytest.mark.parametrize("name", ["rp", "maooam36"])
def test_tgls_wrapper_shape_rules(name):
    from qgs_b200.integrators.integrate import integrate_runge_kutta_tgls
    f, Df, g, _ = setup(name)
    tic = g["ic"][:g["tg3_ic"].shape[0]]
    for tag in ("wt_none", "wt_1d", "wt_2d_ens", "wt_2d_per", "wt_3d"):
        tg = g[tag + "_tg"] if tag + "_tg" in g.files else None
        tt, x, fm = integrate_runge_kutta_tgls(f, Df, 0., 0.5, 0.1, ic=tic, tg_ic=tg, write_steps=2)
        assert np.array_equal(tt, g[tag + "_t"]), tag
        assert rel(x, g[tag + "_x"]) < 1e-10, tag
        assert rel(fm, g[tag + "_fm"]) < 1e-10, tag


# ---- a13: Benettin ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rp", "maooam36"])
@pytest.mark.parametrize("tag", ["full", "sub"])
def test_benettin_vs_golden(name, tag):
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, g, _ = setup(name)
    b, c, a = rk4()
    nv, mdt, ws = g["lyap_%s_meta" % tag]
    nv, ws = int(nv), int(ws)
    lic = g["ic"][:2]
    q0, r0 = g["blv_%s_q0" % tag], g["blv_%s_r0" % tag]
    pre, tim = g["lyap_pretime"], g["lyap_time"]
    rt, re, rv = benettin(f, Df, lic, 0, nv, q0, r0, pre, tim, mdt, ws, False, 1., b, c, a)
    assert rel(rt, g["blv_%s_traj" % tag]) < 1e-10
    assert rel(re, g["blv_%s_exp" % tag]) < 1e-8
    assert rel(rv, g["blv_%s_vec" % tag]) < 1e-8
    rt, re, rv = benettin(f, Df, lic, 1, nv, q0, r0, tim[::-1].copy(), pre[::-1].copy(), mdt, ws, False, 1., b, c, a)
    assert rel(rt, g["flv_%s_traj" % tag]) < 1e-10
    assert rel(re, g["flv_%s_exp" % tag]) < 1e-8
    assert rel(rv, g["flv_%s_vec" % tag]) < 1e-8
