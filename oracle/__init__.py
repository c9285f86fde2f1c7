"""ctypes front-end of the CPU oracle (``oracle/qgs_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``qgs_oracle.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this package; nothing under
``qgs_b200/`` does.  Parity status: pinned against outputs of the unmodified reference
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle.py``).

Function names and argument order follow the reference's jitted functions so the parity tests
read like calls into the reference:

* ``f`` / ``Df``                               <- qgs/functions/tendencies.py:98-121
* ``integrate_runge_kutta_jit``               <- qgs/integrators/integrate.py:182-223
* ``integrate_runge_kutta_tgls_jit``          <- qgs/integrators/integrate.py:555-614
* ``compute_backward_lyap`` / ``compute_forward_lyap`` <- qgs/toolbox/lyapunov.py:471-632
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqgs_oracle.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_long)


def build(force=False):
    """Compile the C restatement with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "qgs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _CTensor(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int), ("rank", ctypes.c_int), ("nnz", ctypes.c_long),
                ("coo", _ip), ("val", _dp), ("jnnz", ctypes.c_long), ("jcoo", _ip), ("jval", _dp)]


class _CPlan(ctypes.Structure):
    _fields_ = [("n_pre", ctypes.c_long), ("n_rec_steps", ctypes.c_long), ("start_idx", _lp),
                ("sub_ptr", _lp), ("sub_time", _dp), ("dt", _dp), ("final_idx", ctypes.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.qgso_n_records.restype = ctypes.c_long
        _lib.qgso_n_records.argtypes = [ctypes.c_long, ctypes.c_long]
        _lib.qgso_num_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_num_threads(n):
    lib().qgso_set_num_threads(ctypes.c_int(int(n)))


def num_threads():
    return int(lib().qgso_num_threads())


class Tensor:
    """The four arrays ``create_tendencies`` extracts (tendencies.py:92-96)."""

    def __init__(self, ndim, coo, val, jcoo=None, jval=None):
        self.ndim = int(ndim)
        self.coo = np.ascontiguousarray(coo, dtype=np.int32)
        self.val = _f64(val)
        self.rank = int(self.coo.shape[1])
        if jcoo is None:
            jcoo = np.zeros((0, self.rank), dtype=np.int32)
            jval = np.zeros((0,))
        self.jcoo = np.ascontiguousarray(jcoo, dtype=np.int32)
        self.jval = _f64(jval)
        self.c = _CTensor(self.ndim, self.rank, len(self.val), self.coo.ctypes.data_as(_ip), _d(self.val),
                          len(self.jval), self.jcoo.ctypes.data_as(_ip), _d(self.jval))

    @classmethod
    def from_npz(cls, path):
        z = np.load(path)
        return cls(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])


def n_records(L, write_steps):
    return int(lib().qgso_n_records(int(L), int(write_steps)))


def sparse_mul3(coo, value, vec_a, vec_b):
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    value, vec_a, vec_b = _f64(value), _f64(vec_a), _f64(vec_b)
    res = np.empty_like(vec_a)
    lib().qgso_sparse_mul3(ctypes.c_long(len(value)), coo.ctypes.data_as(_ip), _d(value), ctypes.c_int(len(vec_a)),
                           _d(vec_a), _d(vec_b), _d(res))
    return res


def sparse_mul2(coo, value, vec):
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    value, vec = _f64(value), _f64(vec)
    res = np.empty((len(vec), len(vec)))
    lib().qgso_sparse_mul2(ctypes.c_long(len(value)), coo.ctypes.data_as(_ip), _d(value), ctypes.c_int(len(vec)),
                           _d(vec), _d(res))
    return res


def sparse_mul5(coo, value, va, vb, vc, vd):
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    value, va, vb, vc, vd = map(_f64, (value, va, vb, vc, vd))
    res = np.empty_like(va)
    lib().qgso_sparse_mul5(ctypes.c_long(len(value)), coo.ctypes.data_as(_ip), _d(value), ctypes.c_int(len(va)),
                           _d(va), _d(vb), _d(vc), _d(vd), _d(res))
    return res


def sparse_mul4(coo, value, va, vb, vc):
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    value, va, vb, vc = map(_f64, (value, va, vb, vc))
    res = np.empty((len(va), len(va)))
    lib().qgso_sparse_mul4(ctypes.c_long(len(value)), coo.ctypes.data_as(_ip), _d(value), ctypes.c_int(len(va)),
                           _d(va), _d(vb), _d(vc), _d(res))
    return res


def f(T, x):
    """Batched tendencies: x (n,) or (N, n)."""
    x2 = _f64(np.atleast_2d(x))
    out = np.empty_like(x2)
    lib().qgso_tendencies(ctypes.byref(T.c), ctypes.c_long(x2.shape[0]), _d(x2), _d(out))
    return out.reshape(np.shape(x))


def Df(T, x):
    """Batched Jacobian: x (n,) -> (n, n); (N, n) -> (N, n, n)."""
    x2 = _f64(np.atleast_2d(x))
    n = T.ndim
    out = np.empty((x2.shape[0], n, n))
    lib().qgso_jacobian(ctypes.byref(T.c), ctypes.c_long(x2.shape[0]), _d(x2), _d(out))
    return out[0] if np.ndim(x) == 1 else out


def rk4_tableau():
    c = np.array([0., 0.5, 0.5, 1.])
    b = np.array([1. / 6, 1. / 3, 1. / 3, 1. / 6])
    a = np.zeros((4, 4))
    a[1, 0] = 0.5
    a[2, 1] = 0.5
    a[3, 2] = 1.
    return b, c, a


def integrate_runge_kutta_jit(T, time, ic, time_direction, write_steps, b, c, a):
    time, ic, b, c, a = map(_f64, (time, ic, b, c, a))
    N, n = ic.shape
    R = n_records(len(time), write_steps)
    traj = np.empty((N, n, R))
    lib().qgso_rk_integrate(ctypes.byref(T.c), ctypes.c_long(N), _d(ic), ctypes.c_long(len(time)), _d(time),
                            ctypes.c_int(int(time_direction)), ctypes.c_long(int(write_steps)),
                            ctypes.c_int(len(b)), _d(a), _d(b), _d(c), _d(traj))
    return traj


def integrate_runge_kutta_tgls_jit(T, time, ic, tg_ic, time_direction, write_steps, b, c, a, adjoint, inverse):
    time, ic, tg_ic, b, c, a = map(_f64, (time, ic, tg_ic, b, c, a))
    N, n = ic.shape
    m = tg_ic.shape[2]
    R = n_records(len(time), write_steps)
    traj = np.empty((N, n, R))
    fm = np.empty((N, n, m, R))
    lib().qgso_rk_tgls_integrate(ctypes.byref(T.c), ctypes.c_long(N), _d(ic), ctypes.c_int(m), _d(tg_ic),
                                 ctypes.c_long(len(time)), _d(time), ctypes.c_int(int(time_direction)),
                                 ctypes.c_long(int(write_steps)), ctypes.c_int(len(b)), _d(a), _d(b), _d(c),
                                 ctypes.c_int(1 if adjoint else 0), ctypes.c_double(float(inverse)),
                                 _d(traj), _d(fm))
    return traj, fm


def qr(A):
    A = _f64(A)
    n, m = A.shape
    Q = np.empty((n, m))
    R = np.empty((m, m))
    work = np.empty(n * m + m)
    lib().qgso_qr(ctypes.c_int(n), ctypes.c_int(m), _d(A), _d(Q), _d(R), _d(work))
    return Q, R


def _lyap(T, ttraj, start_idx, subs, dts, n_pre, final_idx, forward, n_vec, write_steps, n_rec_time_len,
          adjoint, inverse, b, c, a, q0, r0):
    N, n, Ltot = ttraj.shape
    sub_ptr = np.zeros(len(subs) + 1, dtype=np.int64)
    sub_ptr[1:] = np.cumsum([len(s_) for s_ in subs])
    sub_time = _f64(np.concatenate(subs)) if subs else np.zeros(0)
    start_idx = np.ascontiguousarray(start_idx, dtype=np.int64)
    dts = _f64(dts)
    plan = _CPlan(int(n_pre), int(len(subs) - n_pre), start_idx.ctypes.data_as(_lp), sub_ptr.ctypes.data_as(_lp),
                  _d(sub_time), _d(dts), int(final_idx))
    R = n_records(n_rec_time_len, write_steps)
    m = int(n_vec)
    rec_traj = np.empty((N, n, R))
    rec_exp = np.empty((N, m, R))
    rec_vec = np.empty((N, n, m, R))
    q0 = _f64(q0)
    b, c, a = map(_f64, (b, c, a))
    r0p = _d(_f64(r0)) if r0 is not None else None
    ttraj = _f64(ttraj)
    lib().qgso_lyap_benettin(ctypes.byref(T.c), ctypes.c_long(N), ctypes.c_long(Ltot), _d(ttraj),
                             ctypes.byref(plan), ctypes.c_int(1 if forward else 0), ctypes.c_int(m), _d(q0), r0p,
                             ctypes.c_long(int(write_steps)), ctypes.c_int(len(b)), _d(a), _d(b),
                             ctypes.c_int(1 if adjoint else 0), ctypes.c_double(float(inverse)),
                             ctypes.c_long(R), _d(rec_traj), _d(rec_exp), _d(rec_vec))
    return rec_traj, rec_exp, rec_vec


def compute_backward_lyap(T, pretime, time, mdt, ic, n_vec, write_steps, adjoint, inverse, b, c, a, q0, r0=None):
    """lyapunov.py:555-632.  q0 (N, n, n_vec) replaces the reference's unseeded random start."""
    pretime, time = _f64(pretime), _f64(time)
    ttraj = integrate_runge_kutta_jit(T, np.concatenate((pretime[:-1], time)), ic, 1, 1, b, c, a)   # :558
    lp = len(pretime)
    subs, start, dts = [], [], []
    for ti, (tt, dt) in enumerate(zip(pretime[:-1], np.diff(pretime))):                                 # :597
        subs.append(np.concatenate((np.arange(tt, tt + dt, mdt), np.full((1,), tt + dt))))             # :598
        start.append(ti)
        dts.append(dt)
    n_pre = len(subs)
    for ti, (tt, dt) in enumerate(zip(time[:-1], np.diff(time))):                                       # :609
        subs.append(np.concatenate((np.arange(tt, tt + dt, mdt), np.full((1,), tt + dt))))             # :619
        start.append(lp - 1 + ti)
        dts.append(dt)
    final_idx = lp - 1 + len(time) - 1
    return _lyap(T, ttraj, start, subs, dts, n_pre, final_idx, False, n_vec, write_steps, len(time),
                 adjoint, inverse, b, c, a, q0, r0)


def compute_forward_lyap(T, time, posttime, mdt, ic, n_vec, write_steps, adjoint, inverse, b, c, a, q0, r0=None):
    """lyapunov.py:471-552."""
    time, posttime = _f64(time), _f64(posttime)
    ttraj = integrate_runge_kutta_jit(T, np.concatenate((time[:-1], posttime)), ic, 1, 1, b, c, a)  # :474
    Ltot = ttraj.shape[2]
    lt = len(time)
    rposttime, rtime = posttime[::-1], time[::-1]
    subs, start, dts = [], [], []
    for ti, (tt, dt) in enumerate(zip(rposttime[:-1], np.diff(rposttime))):                             # :512
        sub = np.concatenate((np.arange(tt + dt, tt, mdt), np.full((1,), tt)))                          # :514
        subs.append(sub[::-1].copy())                                                                   # time_direction -1
        start.append(Ltot - 1 - ti)
        dts.append(dt)
    n_pre = len(subs)
    for ti, (tt, dt) in enumerate(zip(rtime[:-1], np.diff(rtime))):                                     # :526
        sub = np.concatenate((np.arange(tt + dt, tt, mdt), np.full((1,), tt)))                          # :537
        subs.append(sub[::-1].copy())
        start.append(lt - 1 - ti)
        dts.append(dt)
    final_idx = start[-1] if len(start) > n_pre else 0
    return _lyap(T, ttraj, start, subs, dts, n_pre, final_idx, True, n_vec, write_steps, len(time),
                 adjoint, inverse, b, c, a, q0, r0)


# ---- Ginelli backward recursion (numpy restatement) -------------------------------------------------------
def solve_triangular_matrix(a, b):
    """qgs/functions/util.py:78-98: column i of x solves the leading (i+1) x (i+1) block."""
    x = np.zeros_like(a)
    for i in range(2, a.shape[0] + 1):
        x[:i, i - 1] = np.linalg.solve(a[:i, :i], b[:i, i - 1])
    x[0, 0] = b[0, 0] / a[0, 0]
    return x


def normalize_matrix_columns(a):
    """qgs/functions/util.py:56-75."""
    an = np.zeros_like(a)
    norm = np.zeros(a.shape[0])
    for i in range(a.shape[1]):
        norm[i] = np.linalg.norm(a[:, i], 2)
        an[:, i] = a[:, i] / norm[i]
    return an, norm


def clv_ginelli_backward(tmp_traj, tmp_vec, tmp_R, am, noise, noise_pert, tw, tew, write_steps, dte, n_records):
    """Parts four and five of ``_compute_clv_gin_jit`` (qgs/toolbox/lyapunov.py:1252-1286) for ONE trajectory,
    with the random start matrix ``am`` and the per-step diagonal ``noise (tew, n_vec)`` passed in instead of
    drawn from numba's generator.  ``tmp_traj (tw+1, n)``, ``tmp_vec (tw+1, n, m)``, ``tmp_R (tew, m, m)``."""
    n_dim, n_vec = tmp_vec.shape[1], tmp_vec.shape[2]
    recorded_traj = np.zeros((n_dim, n_records))
    recorded_exp = np.zeros((n_vec, n_records))
    recorded_vec = np.zeros((n_dim, n_vec, n_records))
    for ti in range(tew - 1, tw, -1):
        am_new = solve_triangular_matrix(tmp_R[ti], am)
        for i in range(n_vec):
            am_new[i, i] += (noise[ti, i] if noise is not None else 0.) * noise_pert
        am, norm = normalize_matrix_columns(am_new)
    iw = 1
    mloc_exp = np.ones(n_vec)
    for ti in range(tw, -1, -1):
        am_new = solve_triangular_matrix(tmp_R[ti], am)
        for i in range(n_vec):
            am_new[i, i] += (noise[ti, i] if noise is not None else 0.) * noise_pert
        am, mloc_exp = normalize_matrix_columns(am_new)
        if write_steps > 0 and np.mod(tw - ti, write_steps) == 0:
            recorded_traj[:, -iw] = tmp_traj[ti]
            recorded_exp[:, -iw] = -np.log(np.abs(mloc_exp)) / dte[ti]
            recorded_vec[:, :, -iw] = tmp_vec[ti] @ am
            iw += 1
    recorded_traj[:, 0] = tmp_traj[0]
    recorded_exp[:, 0] = -np.log(np.abs(mloc_exp)) / dte[0]
    recorded_vec[:, :, 0] = tmp_vec[0] @ am
    return recorded_traj, recorded_exp, recorded_vec
