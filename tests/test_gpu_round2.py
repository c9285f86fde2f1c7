"""GPU tests of the round-2 work: the Householder and Cholesky QR of the Benettin kernel, start bases drawn on the device,
member batches bounded by device memory, several devices behind one call, and the long-run statistical criterion of
BASELINE.json on the model the metric is quoted on (MAOOAM-36).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
_cache = {}


def model(name):
    if name not in _cache:
        import oracle
        from qgs_b200.functions.tendencies import tendencies_from_tensor
        z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
        T = oracle.Tensor.from_npz(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        _cache[name] = (f, Df, T)
    return _cache[name]


class env(object):
    """Set environment variables for the duration of a block (libqgsb reads its A/B switches with getenv per launch)."""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _benettin(name, N, n_vec, q0, r0, seed=None, vectors=True, mode=0, mdt=0.1, t1=2.5):
    import oracle
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, T = model(name)
    n = f.ndim
    b, c, a = oracle.rk4_tableau()
    ic = np.random.default_rng(5).random((N, n)) * 0.01
    pre = np.concatenate((np.arange(0., 0.5, 0.1), [0.5]))
    tim = np.concatenate((np.arange(0.5, t1, 0.1), [t1]))
    return benettin(f, Df, ic, mode, n_vec, q0, r0, pre, tim, mdt, 3, False, 1., b, c, a, want_vectors=vectors, seed=seed)


# ---- Cholesky QR on the steps nobody looks at --------------------------------------------------------------------------
@pytest.mark.parametrize("name,n_vec", [("maooam36", 36), ("maooam36", 24), ("maooam36", 17), ("maooam36", 10),
                                        ("maooam36", 9), ("rp", 20), ("rp", 12), ("dynT", 38), ("dynT", 30)])
@pytest.mark.parametrize("vectors", [True, False])
def test_cholesky_qr_steps_reproduce_the_householder_run(name, n_vec, vectors):
    """Between two records the Benettin kernel factorises with a Cholesky QR on the FP64 tensor cores (Gram matrix by
    mma.sync, warp-level Cholesky, forward substitution: pack::chol_factor / chol_solve) and keeps Householder -- the
    reference's np.linalg.qr, lyapunov.py:602-604 -- for every step whose Q or R is recorded or returned.  Householder's Q
    does not depend on the signs of the columns it is given, so the RECORDED vectors keep np.linalg.qr's signs; the
    run with QGSB_QR_CHOL=0 (Householder everywhere, what the golden tests pin against the reference) must be
    reproduced to rounding: exponents (log|diag R|) and vectors, with and without vector records, the full basis and
    partial bases in each compile-time column capacity (16, 24, 32), odd vector counts included."""
    f, Df, T = model(name)
    n = f.ndim
    N = 23
    rng = np.random.default_rng(11)
    q0 = np.stack([np.linalg.qr(rng.random((n, n_vec)))[0] for _ in range(N)])
    out = {}
    for chol in ("0", "1"):
        with env(QGSB_QR_CHOL=chol):
            out[chol] = _benettin(name, N, n_vec, q0, None, vectors=vectors, t1=6.5)
    assert np.array_equal(out["0"][0], out["1"][0])                       # the trajectory never sees the basis
    assert np.abs(out["0"][1] - out["1"][1]).max() < 1e-11 * max(1., np.abs(out["0"][1]).max())
    if vectors:
        assert np.abs(out["0"][2] - out["1"][2]).max() < 1e-11
        q = out["1"][2][..., -1]
        assert np.abs(np.einsum("nij,nik->njk", q, q) - np.eye(n_vec)).max() < 1e-13


@pytest.mark.parametrize("mode,adjoint", [(1, False), (0, True), (1, True)])
def test_cholesky_qr_steps_in_the_forward_vector_and_adjoint_modes(mode, adjoint):
    """The same equivalence for the other drivers of the Benettin kernel: the forward-Lyapunov-vector pass (mode 1: stored
    trajectory, time walked backwards, lyapunov.py:471-550) and the adjoint tangent model."""
    import oracle
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, T = model("maooam36")
    n = f.ndim
    N = 11
    b, c, a = oracle.rk4_tableau()
    rng = np.random.default_rng(13)
    ic = rng.random((N, n)) * 0.01
    q0 = np.stack([np.linalg.qr(rng.random((n, n)))[0] for _ in range(N)])
    pre = np.concatenate((np.arange(0., 0.5, 0.1), [0.5]))
    tim = np.concatenate((np.arange(0.5, 4.5, 0.1), [4.5]))
    if mode == 1:
        pre, tim = tim[::-1].copy(), pre[::-1].copy()
    out = {}
    for chol in ("0", "1"):
        with env(QGSB_QR_CHOL=chol):
            out[chol] = benettin(f, Df, ic, mode, n, q0, None, pre, tim, 0.1, 3, adjoint, 1., b, c, a)
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.abs(out["0"][1] - out["1"][1]).max() < 1e-11
    assert np.abs(out["0"][2] - out["1"][2]).max() < 1e-11


def test_cholesky_qr_falls_back_to_householder_on_an_ill_conditioned_step():
    """Orthogonality of a Cholesky QR degrades with cond(A)^2, and the Gram matrix of a rank-deficient basis has no
    Cholesky factor at all.  A start basis whose second column repeats the first, and whose fourth is the third up to
    1e-9, makes the first propagated matrix singular to working precision: the pivot test of chol_factor must refuse,
    the block then takes the Householder code for that step (whose arithmetic is exactly the Householder-only run's),
    and everything stays finite and orthonormal.  Later steps are well conditioned again and differ from the
    Householder-only run by rounding only."""
    import oracle
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, T = model("maooam36")
    n = f.ndim
    N = 9
    b, c, a = oracle.rk4_tableau()
    rng = np.random.default_rng(3)
    ic = rng.random((N, n)) * 0.01
    q0 = np.stack([np.linalg.qr(rng.random((n, n)))[0] for _ in range(N)])
    q0[:, :, 1] = q0[:, :, 0]
    q0[:, :, 3] = q0[:, :, 2] + 1e-9 * q0[:, :, 5]
    pre = np.concatenate((np.arange(0., 0.5, 0.1), [0.5]))
    tim = np.concatenate((np.arange(0.5, 3.5, 0.1), [3.5]))
    out = {}
    for chol in ("0", "1"):
        with env(QGSB_QR_CHOL=chol):
            out[chol] = benettin(f, Df, ic, 0, n, q0, None, pre, tim, 0.1, 4, False, 1., b, c, a)
    for x, y in zip(out["0"], out["1"]):
        assert np.all(np.isfinite(x)) and np.all(np.isfinite(y))
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.abs(out["0"][1] - out["1"][1]).max() < 1e-9
    assert np.abs(out["0"][2] - out["1"][2]).max() < 1e-9
    q = out["1"][2][..., -1]
    assert np.abs(np.einsum("nij,nik->njk", q, q) - np.eye(n)).max() < 1e-12


def test_a_refused_cholesky_step_does_not_touch_the_other_members_of_the_block():
    """The pivot test is per member: a member with a singular basis goes through the Householder code for that step, the
    members that share its thread block keep their Cholesky QR -- their results are BITWISE what they are when the
    singular member is not there.  (Members sharded over devices or cut into batches must stay bitwise one launch.)"""
    import oracle
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, T = model("maooam36")
    n = f.ndim
    N = 9
    b, c, a = oracle.rk4_tableau()
    rng = np.random.default_rng(4)
    ic = rng.random((N, n)) * 0.01
    q0 = np.stack([np.linalg.qr(rng.random((n, n)))[0] for _ in range(N)])
    bad = q0.copy()
    bad[3, :, 7] = bad[3, :, 6]
    pre = np.concatenate((np.arange(0., 0.5, 0.1), [0.5]))
    tim = np.concatenate((np.arange(0.5, 2.5, 0.1), [2.5]))
    good = benettin(f, Df, ic, 0, n, q0, None, pre, tim, 0.1, 4, False, 1., b, c, a)
    mixed = benettin(f, Df, ic, 0, n, bad, None, pre, tim, 0.1, 4, False, 1., b, c, a)
    others = [g for g in range(N) if g != 3]
    for x, y in zip(good, mixed):
        assert np.all(np.isfinite(y))
        assert np.array_equal(x[others], y[others])
    assert not np.array_equal(good[2][3], mixed[2][3])


# ---- the re-orthonormalisations are the same arithmetic -----------------------------------------------------------------
@pytest.mark.parametrize("name,n_vec", [("maooam36", 36), ("maooam36", 10), ("rp", 20), ("dynT", 38)])
def test_rolled_and_unrolled_qr_are_bitwise_equal(name, n_vec):
    """The unrolled Householder factorisation (one code block per reflector) and the rolled one (three reflector groups,
    rows <= j multiplied by an explicit zero) differ in instruction bytes, never in a value; neither does the way the
    columns are dealt to the threads (QGSB_QR_REMAP).  Trajectories, exponents and vectors must be IDENTICAL.  (The
    defaults are what the golden tests of test_gpu_parity.py compare with the reference's np.linalg.qr.)"""
    f, Df, T = model(name)
    n = f.ndim
    N = 23                                     # not a multiple of the members per block: partially filled last block
    rng = np.random.default_rng(8)
    q0 = np.stack([np.linalg.qr(rng.random((n, n_vec)))[0] for _ in range(N)])
    out = []
    for rolled, remap in (("0", "1"), ("1", "1"), ("0", "0"), ("1", "0")):
        with env(QGSB_QR_ROLLED=rolled, QGSB_QR_REMAP=remap):
            out.append(_benettin(name, N, n_vec, q0, None))
    for other in out[1:]:
        for x, y in zip(out[0], other):
            assert np.array_equal(x, y)


# ---- start bases drawn on the device -------------------------------------------------------------------------------------
def test_device_start_bases_are_reproducible_orthonormal_and_member_indexed():
    """q0 = NULL: qr(random((n_dim, n_vec))) of lyapunov.py:592-593 on the device.  Same seed -> same run; another seed
    -> another basis; the recorded vectors are orthonormal; a member's draw depends on its GLOBAL index only, so
    cutting the ensemble into member batches (forced here by a tiny memory budget) changes nothing."""
    N, n_vec = 40, 12
    a1 = _benettin("maooam36", N, n_vec, None, None, seed=1234)
    a2 = _benettin("maooam36", N, n_vec, None, None, seed=1234)
    b1 = _benettin("maooam36", N, n_vec, None, None, seed=99)
    for x, y in zip(a1, a2):
        assert np.array_equal(x, y)
    assert not np.array_equal(a1[2], b1[2]) and np.array_equal(a1[0], b1[0])      # vectors differ, trajectories do not
    q = np.moveaxis(a1[2], 3, 1)                                                  # (N, R, n, m)
    gram = np.einsum('mrik,mril->mrkl', q, q)
    assert np.max(np.abs(gram - np.eye(n_vec))) < 1e-12
    # members are distinct draws
    assert np.max(np.abs(a1[2][0] - a1[2][1])) > 1e-3
    with env(QGSB_TANGENT_BUDGET_MB="1"):                                          # a handful of members per batch
        c1 = _benettin("maooam36", N, n_vec, None, None, seed=1234)
    for x, y in zip(a1, c1):
        assert np.array_equal(x, y)
    # the generic (block per member) kernels draw the same numbers and factorise them to the same basis
    with env(QGSB_TGLS_KERNEL="generic"):
        g1 = _benettin("maooam36", N, n_vec, None, None, seed=1234)
    assert np.max(np.abs(g1[2] - a1[2])) < 1e-9 and np.max(np.abs(g1[1][:, :, 1:] - a1[1][:, :, 1:])) < 1e-9


def test_first_exponent_record_of_a_device_drawn_basis_matches_numpy_qr():
    """With no convergence phase the first recorded exponents are log|diag R| / dt of the QR of the DRAWN matrix
    (lyapunov.py:592-593, :611): recover the drawn matrix from Q and R?  Not available -- instead check the
    invariant numpy gives for a uniform [0, 1) matrix: |R_00| = norm of the first column, in (0, sqrt(n)), and the
    recorded vectors of record 0 are the orthonormal start basis itself."""
    import oracle
    from qgs_b200.toolbox.lyapunov import benettin
    f, Df, T = model("maooam36")
    b, c, a = oracle.rk4_tableau()
    ic = np.random.default_rng(5).random((16, 36)) * 0.01
    pre = np.array([0.])                                   # t0 == tw: no convergence step
    tim = np.concatenate((np.arange(0., 0.5, 0.1), [0.5]))
    traj, exp, vec = benettin(f, Df, ic, 0, 36, None, None, pre, tim, 0.1, 1, False, 1., b, c, a, seed=7)
    r00 = np.exp(exp[:, 0, 0] * 0.1)                       # |R_00| of the start factorisation
    # a column of 36 uniform numbers has norm sqrt(36 / 3) = 3.46 on average, within (1.7, 4.8) with overwhelming odds
    assert np.all((r00 > 1.7) & (r00 < 4.8)), r00
    q = vec[:, :, :, 0]
    assert np.max(np.abs(np.einsum('mik,mil->mkl', q, q) - np.eye(36))) < 1e-12
    # LAPACK's sign convention on a positive matrix: R_00 = -norm, so Q[:, 0] = -column / norm has no positive entry
    assert np.all(q[:, :, 0] <= 0.)


def test_tangent_records_in_member_batches_equal_one_batch():
    """qgsb_rk_tgls_integrate keeps (R, N, n + n m) records on the device; a long write_steps = 1 run that does not fit
    is cut into member batches.  Forced here with a 1 MB budget: results are identical."""
    from qgs_b200.integrators.integrator import RungeKuttaTglsIntegrator
    f, Df, T = model("maooam36")
    rng = np.random.default_rng(2)
    ic = rng.random((50, 36)) * 0.01
    out = []
    for budget in (None, "1"):
        with env(QGSB_TANGENT_BUDGET_MB=budget):
            integ = RungeKuttaTglsIntegrator()
            integ.set_func(f, Df)
            integ.integrate(0., 1., 0.1, ic=ic, write_steps=1)
            out.append(integ.get_trajectories())
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


# ---- several devices behind one call ---------------------------------------------------------------------------------------
def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_devices() < 2")
def test_members_sharded_over_the_devices_of_one_process_are_bitwise_one_device():
    """qgsb_rk_integrate / qgsb_rk_tgls_integrate / qgsb_lyap_benettin / qgsb_clv_ginelli split the members over every
    device the process drives (the reference deals trajectories to num_threads workers, integrator.py:386-395).  Same results, bit for
    bit, as on one device -- through the unchanged Python classes."""
    from qgs_b200 import _lib
    from qgs_b200.integrators.integrator import RungeKuttaIntegrator, RungeKuttaTglsIntegrator
    from qgs_b200.toolbox.lyapunov import LyapunovsEstimator
    f, Df, T = model("maooam36")
    rng = np.random.default_rng(4)
    ic = rng.random((3 * 8192 + 77, 36)) * 0.01
    G = _n_devices()
    res = {}
    try:
        for devs in ([0], list(range(G))):
            _lib.set_devices(devs)
            assert _lib.device_count() == len(devs)
            integ = RungeKuttaIntegrator()
            integ.set_func(f)
            integ.integrate(0., 2., 0.1, ic=ic, write_steps=0)
            end = integ.get_trajectories()[1]
            integ.integrate(0., 1., 0.1, ic=ic, write_steps=3)
            traj = integ.get_trajectories()[1]
            tg = RungeKuttaTglsIntegrator()
            tg.set_func(f, Df)
            tg.integrate(0., 0.5, 0.1, ic=ic[:4096], tg_ic=np.eye(36)[:5], write_steps=0)
            tl = tg.get_trajectories()
            est = LyapunovsEstimator()
            est.set_func(f, Df)
            np.random.seed(5)
            est.compute_lyapunovs(0., 0.5, 1.5, 0.1, 0.1, ic=ic[:4100], write_steps=5, n_vec=36, vectors=False)
            ly = est.get_lyapunovs()
            from qgs_b200.toolbox.lyapunov import CovariantLyapunovsEstimator
            frp, Dfrp, _ = model("rp")
            clv = CovariantLyapunovsEstimator()
            clv.set_func(frp, Dfrp)
            np.random.seed(9)
            clv.compute_clvs(0., 0.5, 1., 1.5, 0.1, 0.1, ic=np.random.default_rng(6).random((1100, 20)) * 0.1,
                             write_steps=2, method=0)
            cv = clv.get_clvs()
            res[len(devs)] = (end, traj, tl[1], tl[2], ly[1], ly[2], cv[1], cv[2], cv[3])
    finally:
        _lib.set_devices([0])
    for x, y in zip(res[1], res[G]):
        assert np.array_equal(x, y)


@pytest.mark.skipif("_n_devices() < 2")
def test_resident_ensemble_spread_over_the_devices_of_one_process():
    """qgsb_ensemble_* with several devices: a large resident ensemble is a composite of one block per device.  States
    and streamed trajectories are bitwise those of one device; the moments agree to rounding (other summation order)."""
    from qgs_b200 import _lib
    from qgs_b200.ensemble import DeviceEnsemble
    f, Df, T = model("maooam36")
    ic = np.random.default_rng(14).random((2 * 8192 + 300, 36)) * 0.01
    G = _n_devices()
    res = {}
    try:
        for devs in ([0], list(range(G))):
            _lib.set_devices(devs)
            ens = DeviceEnsemble(f, ic)
            ens.integrate(0., 1., 0.1)
            a = ens.states()
            t, traj = ens.integrate_trajectories(1., 1.6, 0.1, write_steps=2)
            tm, mean, var = ens.integrate_moments(1.6, 2.2, 0.1, write_steps=3)
            m1, v1 = ens.moments()
            ens.set_states(a)
            b = ens.states()
            ens.close()
            res[len(devs)] = (a, traj, mean, var, m1, v1, b)
    finally:
        _lib.set_devices([0])
    one, many = res[1], res[G]
    assert np.array_equal(one[0], many[0]) and np.array_equal(one[1], many[1]) and np.array_equal(one[6], many[6])
    assert np.array_equal(one[0], one[6])
    for k in (2, 3, 4, 5):
        assert np.allclose(one[k], many[k], rtol=1e-12, atol=1e-18)


# ---- long-run statistics on the headline model ---------------------------------------------------------------------------
def test_maooam36_long_run_moments_and_leading_exponents_match_the_oracle():
    """BASELINE.json's third criterion on MAOOAM-36 itself: an ensemble is spun up ON THE GPU to the attractor
    (qgs_maooam.py:69: 3e6 time units = 3e7 RK4 steps on the latency-regime kernel, about 40 s;
    QGSB_TEST_SPINUP shortens it), then integrated for 1000 time units (~20 Lyapunov times
    of ~50 units) by the GPU and by the CPU oracle from the same states.  After a few Lyapunov times
    the two ensembles are different samples of the same attractor: the time-and-ensemble means of ALL 36 variables must
    agree within 5 standard errors (standard error from the spread of the members' time means, both sides), the
    standard deviations within 10 % + 5 standard errors, and the 6 leading Lyapunov exponents (Benettin, 10 vectors,
    different start bases on the two sides) within 4 standard errors."""
    import oracle
    from qgs_b200.integrators.integrator import RungeKuttaIntegrator
    from qgs_b200.toolbox.lyapunov import LyapunovsEstimator
    f, Df, T = model("maooam36")
    b, c, a = oracle.rk4_tableau()
    rng = np.random.default_rng(2024)
    N = 192
    integ = RungeKuttaIntegrator()
    integ.set_func(f)
    spinup = float(os.environ.get("QGSB_TEST_SPINUP", "3.e6"))
    integ.integrate(0., spinup, 0.1, ic=rng.random((N, 36)) * 0.01, write_steps=0)
    ic = np.ascontiguousarray(integ.get_trajectories()[1])
    assert np.all(np.isfinite(ic)) and np.max(np.abs(ic)) < 1.

    # (a) climatological moments over 1000 time units, records every 10 steps
    integ.integrate(0., 1000., 0.1, ic=ic, write_steps=10)
    tg, g = integ.get_trajectories()
    tv = np.concatenate((np.arange(0., 1000., 0.1), [1000.]))
    o = oracle.integrate_runge_kutta_jit(T, tv, ic, 1, 10, b, c, a)
    assert g.shape == o.shape == (N, 36, 1001)
    # short horizon: trajectory parity proper (2 time units << Lyapunov time)
    assert np.max(np.abs(g[:, :, :3] - o[:, :, :3])) < 1e-10 * np.max(np.abs(o))
    gm, om = g.mean(axis=2), o.mean(axis=2)                       # members' time means (N, 36)
    se = np.sqrt(gm.var(axis=0) / N + om.var(axis=0) / N)
    scale = np.maximum(np.abs(om.mean(axis=0)), om.std(axis=0))
    z = np.abs(gm.mean(axis=0) - om.mean(axis=0)) / np.maximum(se, 1e-12 * scale)
    assert np.all(z < 5.), z
    gs, os_ = g.std(axis=2), o.std(axis=2)                        # members' standard deviations in time
    se_s = np.sqrt(gs.var(axis=0) / N + os_.var(axis=0) / N)
    assert np.all(np.abs(gs.mean(axis=0) - os_.mean(axis=0)) < 0.1 * os_.mean(axis=0) + 5. * se_s)

    # (b) leading exponents: 64 members, 10 vectors, 200 units of convergence + 800 recorded
    M, n_vec = 64, 10
    np.random.seed(11)
    est = LyapunovsEstimator()
    est.set_func(f, Df)
    est.compute_lyapunovs(0., 200., 1000., 0.1, 0.1, ic=ic[:M], write_steps=10, n_vec=n_vec, vectors=False)
    ge = est.get_lyapunovs()[2].reshape(M, n_vec, -1)[:, :, 1:].mean(axis=2)
    pre = np.concatenate((np.arange(0., 200., 0.1), [200.]))
    tim = np.concatenate((np.arange(200., 1000., 0.1), [1000.]))
    q0 = np.stack([np.linalg.qr(rng.random((36, n_vec)))[0] for _ in range(M)])
    oe = oracle.compute_backward_lyap(T, pre, tim, 0.1, ic[:M], n_vec, 10, False, 1., b, c, a, q0,
                                      np.stack([np.eye(n_vec)] * M))[1][:, :, 1:].mean(axis=2)
    se_e = np.sqrt(ge.var(axis=0) / M + oe.var(axis=0) / M)
    ze = np.abs(ge.mean(axis=0) - oe.mean(axis=0)) / np.maximum(se_e, 1e-6)
    assert np.all(ze[:6] < 4.), (ze, ge.mean(axis=0), oe.mean(axis=0))
    assert ge.mean(axis=0)[0] > 0.                                # MAOOAM at these parameters is chaotic


# ---- initialize() on resident device batches ------------------------------------------------------------------------------
def test_initialize_in_device_batches_matches_an_oracle_replay_of_the_same_draws():
    """RungeKuttaIntegrator.initialize (integrator.py:198-295): first batch converged over the long transient, every
    further batch = previous batch + pert_size * randn, reconverged over the short one.  The batch is the device's
    (here forced to 64), num_threads is only a hint.  Replaying the SAME numpy draws through the CPU oracle must give
    the same members (transients short against the Lyapunov time, so the comparison is a trajectory comparison)."""
    import oracle
    from qgs_b200.integrators.integrator import RungeKuttaIntegrator
    f, Df, T = model("maooam36")
    b, c, a = oracle.rk4_tableau()
    integ = RungeKuttaIntegrator(num_threads=3)
    integ.set_func(f)
    integ.device_batch = 64
    N, conv, reconv, pert = 150, 6., 2., 1e-3

    def scaled_randn(*shape):                    # N(0, 1) initial conditions blow MAOOAM up: scale the draws
        return 0.01 * base_randn(*shape)

    base_randn = np.random.randn
    np.random.seed(5)
    np.random.randn = scaled_randn
    try:
        integ.initialize(conv, 0.1, pert_size=pert / 0.01, reconvergence_time=reconv, number_of_trajectories=N)
    finally:
        np.random.randn = base_randn
    got = integ.get_ic()
    assert got.shape == (N, 36) and integ.n_traj == N
    # replay
    np.random.seed(5)
    tv = lambda t1: np.concatenate((np.arange(0., t1, 0.1), [t1]))
    x = oracle.integrate_runge_kutta_jit(T, tv(conv), 0.01 * np.random.randn(64, 36), 1, 0, b, c, a)[:, :, 0]
    ref = np.empty((N, 36))
    ref[:64] = x
    index = 64
    while index < N:
        count = min(64, N - index)
        x = oracle.integrate_runge_kutta_jit(T, tv(reconv), x[:count] + pert * np.random.randn(count, 36), 1, 0,
                                             b, c, a)[:, :, 0]
        ref[index:index + count] = x
        index += count
    assert np.max(np.abs(got - ref)) < 1e-10 * np.max(np.abs(ref))
    # a single member: ic is squeezed like get_trajectories() squeezes it; given ics are converged as they are
    integ.initialize(1., 0.1, ic=ref[0])
    assert integ.get_ic().shape == (36,)
    integ.initialize(1., 0.1, ic=ref[:5])
    one = oracle.integrate_runge_kutta_jit(T, tv(1.), ref[:5], 1, 0, b, c, a)[:, :, 0]
    assert np.max(np.abs(integ.get_ic() - one)) < 1e-10 * np.max(np.abs(one))


# ---- large bases: the tangent product as a dense GEMM on the FP64 tensor cores -------------------------------------------
@pytest.mark.parametrize("name,m,adjoint", [("atm6x6", 16, False), ("atm6x6", 150, False), ("atm6x6", 9, True),
                                            ("maooam36", 10, False), ("maooam36", 36, True), ("dynT", 38, False)])
def test_dense_tensor_core_product_equals_the_sparse_product(name, m, adjoint):
    """Generic tangent kernels: out = J @ X over the sparse position list (QGSB_TGLS_DENSE=0) or as a dense
    (n x n) @ (n x m) product with mma.sync.m8n8k4.f64 (=1; the default from 64 variables on).  Same J, same X: the results
    agree to rounding.  Small models are forced through the dense path too, for the padding (n, m not multiples of 8)."""
    from qgs_b200.integrators.integrator import RungeKuttaTglsIntegrator
    f, Df, T = model(name)
    n = f.ndim
    rng = np.random.default_rng(3)
    ic = rng.random((5, n)) * 0.01
    tg = rng.standard_normal((m, n))
    out = []
    for dense in ("0", "1"):
        with env(QGSB_TGLS_KERNEL="generic", QGSB_TGLS_DENSE=dense):
            integ = RungeKuttaTglsIntegrator()
            integ.set_func(f, Df)
            integ.integrate(0., 0.4, 0.1, ic=ic, tg_ic=tg, write_steps=2, adjoint=adjoint)
            out.append(integ.get_trajectories())
    assert np.array_equal(out[0][1], out[1][1])
    scale = np.max(np.abs(out[0][2]))
    assert np.max(np.abs(out[0][2] - out[1][2])) < 1e-12 * scale


def test_dense_product_in_the_generic_benettin_kernel():
    """The same switch inside the Benettin loop (generic kernel, 228 variables, 40 vectors): exponents and vectors of the
    dense product agree with the sparse one's."""
    f, Df, T = model("atm6x6")
    rng = np.random.default_rng(4)
    q0 = np.stack([np.linalg.qr(rng.random((228, 40)))[0] for _ in range(3)])
    out = []
    for dense in ("0", "1"):
        with env(QGSB_TGLS_DENSE=dense):
            out.append(_benettin("atm6x6", 3, 40, q0, None, t1=1.1))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.max(np.abs(out[0][1][:, :, 1:] - out[1][1][:, :, 1:])) < 1e-10 * np.max(np.abs(out[0][1][:, :, 1:]))
    assert np.max(np.abs(out[0][2] - out[1][2])) < 1e-10
