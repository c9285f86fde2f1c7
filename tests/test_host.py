"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, the host logic of the
Python mirror (time vectors, record counts, tg_ic shape rules, Benettin plans, sharding, statistics
collective on gloo with world_size 2), the code generator, and the loud failure without a device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- the boundary ------------------------------------------------------------------------------------
def test_library_loads_and_exports_every_header_symbol():
    from qgs_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "qgsb.h")).read()
    declared = set(re.findall(r"QGSB_API[^;]*?\b(qgsb_\w+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "libqgsb.so does not export %s" % name
    # the ctypes stub binds exactly the header's functions
    assert declared == set(_lib.exported_names())
    assert b"sm_100a" in lib.qgsb_version()


def test_compute_fails_loudly_without_a_device():
    if has_gpu():
        pytest.skip("a GPU is present")
    from qgs_b200.functions.tendencies import tendencies_from_tensor
    from qgs_b200.functions import sparse_mul
    z = np.load(os.path.join(GOLDEN, "tensor_rp.npz"))
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        sparse_mul.sparse_mul3(z["coo"], z["val"], np.ones(21), np.ones(21))


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "qgs_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "qgs_oracle" not in src, fn


def test_non_tensor_callables_are_rejected():
    from qgs_b200.integrators.integrator import RungeKuttaIntegrator
    from qgs_b200.integrators.integrate import integrate_runge_kutta
    with pytest.raises(TypeError, match="no CPU fallback"):
        RungeKuttaIntegrator().set_func(lambda t, x: -x)
    with pytest.raises(TypeError):
        integrate_runge_kutta(lambda t, x: -x, 0., 1., 0.1, ic=np.zeros(3))
    integ = RungeKuttaIntegrator(num_threads=3)
    assert integ.integrate(0., 1., 0.1) == 0           # "No function to integrate defined!" (integrator.py:340-342)
    integ.terminate()                                   # terminate on a never-started integrator is fine
    assert integ.num_threads == 3 and integ.b.shape == (4,) and integ.a[3, 2] == 1.


# ---- host logic ------------------------------------------------------------------------------------------
def test_record_count_and_time_vector_rules():
    from qgs_b200.integrators.integrate import n_records_of, returned_time, directed_dt
    g = np.load(os.path.join(GOLDEN, "golden_rp.npz"))
    time = np.concatenate((np.arange(0., 1.35, 0.1), [1.35]))
    for tag, (fwd, ws) in {"w_fwd_ws4": (True, 4), "w_bwd_ws4": (False, 4), "w_fwd_ws0": (True, 0),
                           "w_bwd_ws5": (False, 5)}.items():
        tt = returned_time(time, 0., 1.35, fwd, ws)
        assert np.array_equal(np.asarray(tt), g[tag + "_t"]), tag
        assert n_records_of(time, ws) == np.atleast_3d(g[tag + "_x"]).shape[-1] or ws == 0
    for L in range(1, 30):
        t = np.arange(L, dtype=float)
        for ws in range(0, 7):
            exp = 1 if ws == 0 else len(t[::ws]) + (1 if t[::ws][-1] != t[-1] else 0)
            assert n_records_of(t, ws) == exp
    # dt is NOT constant: np.diff of an arange
    t = np.concatenate((np.arange(0., 100., 0.1), [100.]))
    d = directed_dt(t, 1)
    assert len(set(d.tolist())) > 1 and np.array_equal(directed_dt(t, -1), -d[::-1])


def test_tg_ic_shape_rules_match_reference_fixture():
    from qgs_b200.integrators.integrate import normalise_tg_ic, restore_fmatrix_orientation
    g = np.load(os.path.join(GOLDEN, "golden_rp.npz"))
    n, n_traj = 20, g["tg3_ic"].shape[0]
    cases = {"wt_1d": (n_traj, n, 1), "wt_2d_ens": (n_traj, n, 4), "wt_2d_per": (n_traj, n, 1),
             "wt_3d": (n_traj, n, 4)}
    for tag, shape in cases.items():
        tg = g[tag + "_tg"]
        norm = normalise_tg_ic(tg, n_traj, n)
        assert norm.shape == shape, tag
        fm = np.zeros(shape + (4,))
        out = np.squeeze(restore_fmatrix_orientation(fm, tg, n))
        assert out.shape == g[tag + "_fm"].shape, tag
    # tg_ic=None with n_traj == n_dim is read as one vector per member (SURVEY.md a12 ambiguity)
    assert normalise_tg_ic(np.eye(5), 5, 5).shape == (5, 5, 1)


def test_benettin_plan_reproduces_numpy_arange_rounding():
    from qgs_b200.toolbox.lyapunov import _subtimes
    pre = np.concatenate((np.arange(0., 1., 0.1), [1.]))
    ptr, sub = _subtimes(pre, 0.05)
    assert ptr[0] == 0 and ptr[-1] == len(sub) and len(ptr) == len(pre)
    k = 0
    for tt, dt in zip(pre[:-1], np.diff(pre)):
        ref = np.diff(np.concatenate((np.arange(tt, tt + dt, 0.05), [tt + dt])))
        assert np.array_equal(sub[ptr[k]:ptr[k + 1]], ref)
        k += 1
    rpre = pre[::-1]
    ptr, sub = _subtimes(rpre, 0.05, backward=True)
    assert np.all(sub <= 0.) and abs(sub.sum() + 1.) < 1e-12


def test_util_helpers():
    from qgs_b200.functions.util import reverse, normalize_matrix_columns, solve_triangular_matrix
    a = np.arange(5.)
    assert np.array_equal(reverse(a), a[::-1])
    rng = np.random.default_rng(0)
    m = rng.standard_normal((6, 6))
    an, norm = normalize_matrix_columns(m)
    assert np.allclose(np.linalg.norm(an, axis=0), 1.) and np.allclose(norm, np.linalg.norm(m, axis=0))
    r = np.triu(rng.standard_normal((6, 6))) + 3 * np.eye(6)
    bmat = np.triu(rng.standard_normal((6, 6)))
    x = solve_triangular_matrix(r, bmat)
    assert np.allclose(np.triu(r @ x), bmat)


def test_jacobian_tensor_from_coo_matches_reference_jacobian():
    from qgs_b200.functions.tendencies import jacobian_tensor_from_coo
    for name in ("rp", "maooam36", "dynT"):
        z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        # the reference derives the Jacobian tensor from the UN-simplified tensor (qgtensor.py:665); from the
        # upper-triangularised one the contraction J_ij = sum_k J_ijk x_k must still agree
        jc, jv = jacobian_tensor_from_coo(z["coo"].astype(int), z["val"])
        n1 = int(z["ndim"]) + 1
        rng = np.random.default_rng(1)
        x = rng.standard_normal(n1)
        x[0] = 1.

        def contract(coo, val):
            out = np.zeros((n1, n1))
            for c, v in zip(coo, val):
                out[c[0], c[1]] += v * np.prod(x[c[2:]])
            return out[1:, 1:]
        assert np.allclose(contract(jc, jv), contract(z["jcoo"].astype(int), z["jval"]), rtol=1e-11, atol=1e-13)


# ---- code generator ----------------------------------------------------------------------------------------
def test_codegen_hash_matches_library_and_modules_exist():
    from qgs_b200 import _lib, codegen
    lib = _lib.load()
    for name in ("rp", "maooam36", "dynT"):
        z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        coo, val = _lib.i32(z["coo"]), _lib.f64(z["val"])
        h = ctypes.c_uint64()
        _lib.check(lib.qgsb_tensor_hash(int(z["ndim"]), int(z["rank"]), len(val), coo.ctypes.data_as(_lib.c_int32_p),
                                        _lib.dptr(val), ctypes.byref(h)))
        cs, vs = codegen.sort_by_row(coo, val)
        assert h.value == codegen.tensor_hash(int(z["ndim"]), int(z["rank"]), cs, vs)
        src, h2, n_fp = codegen.emit_source(name, int(z["ndim"]), int(z["rank"]), cs, vs)
        assert h2 == h.value and ("0x%016xULL" % h2) in src
        gen = os.path.join(REPO, "qgs_b200", "csrc", "generated", "spec_%s.cu" % name)
        assert os.path.exists(gen) and ("0x%016xULL" % h2) in open(gen).read()
    # MAOOAM-36: coefficient factoring brings 559 naive FP64 instructions per f down to < 480
    z = np.load(os.path.join(GOLDEN, "tensor_maooam36.npz"))
    cs, vs = codegen.sort_by_row(z["coo"], z["val"])
    assert codegen.emit_source("m", 36, 3, cs, vs)[2] < 480


def _eval_generated_body(src, x):
    """Run the generated straight-line F_BODY with Python floats (fma = a*b+c)."""
    start = src.index("#define F_BODY")
    lines = []
    for ln in src[start:].splitlines()[1:]:
        if not ln.rstrip().endswith("\\") and not ln.strip():
            break
        lines.append(ln.rstrip().rstrip("\\"))
        if not ln.rstrip().endswith("\\"):
            break
    text = " ".join(lines)
    text = re.sub(r"\((-?0x[0-9a-f.]+p[-+]\d+)\)", lambda m: repr(float.fromhex(m.group(1))), text)
    text = re.sub(r"K\((\d+)\)", r"K[\1]", text)
    text = text.replace("{", ";").replace("}", ";")
    env = {"x": x, "fma": lambda a, b, c: a * b + c, "K": {}, "t": 0., "p2": 0., "p3": 0., "p4": 0.}
    for stmt in text.split(";"):
        stmt = stmt.strip()
        if not stmt or stmt.startswith("ROW_DONE") or stmt.startswith("(void)"):
            continue
        if stmt.startswith("double "):
            stmt = stmt[len("double "):]
            if "=" not in stmt:
                continue
            for piece in stmt.split(","):
                exec(piece.strip(), env)
            continue
        exec(stmt, env)
    return env["K"]


def test_generated_row_code_is_the_same_polynomial():
    """Evaluate the generated straight-line code with Python floats against the plain COO loop (row mode for the
    rank-3 tensors and dynT, monomial-sharing mode for T4)."""
    from qgs_b200 import codegen
    import math
    for name in ("rp", "maooam36", "dynT", "T4"):
        z = np.load(os.path.join(GOLDEN, "tensor_%s.npz" % name))
        n = int(z["ndim"])
        cs, vs = codegen.sort_by_row(z["coo"], z["val"])
        src, _, _ = codegen.emit_source(name, n, int(z["rank"]), cs, vs)
        rng = np.random.default_rng(5)
        x = [1.] + list(rng.standard_normal(n))
        K = _eval_generated_body(src, x)
        ref = np.zeros(n + 1)
        mag = np.zeros(n + 1)
        for c, v in zip(cs, vs):
            term = v * math.prod(x[j] for j in c[1:])
            ref[c[0]] += term
            mag[c[0]] += abs(term)
        for i in range(1, n + 1):
            assert abs(K[i] - ref[i]) <= 1e-12 * max(1., mag[i]), (name, i, K[i], ref[i])


# ---- sharding and the statistics collective (gloo, world_size 2) ---------------------------------------------
def test_shard_bounds_partition():
    from qgs_b200.ensemble import shard_bounds
    for N in (1, 7, 128, 1000, 1 << 20):
        for G in (1, 2, 3, 8):
            parts = [shard_bounds(N, G, g) for g in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == N
            assert all(parts[i][1] == parts[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == -(-N // G)
            assert N < G or min(sizes) >= 1          # no rank is left empty (N = 9, G = 8 used to give 2,2,2,2,1,0,0,0)
    assert [shard_bounds(9, 8, g)[1] - shard_bounds(9, 8, g)[0] for g in range(8)].count(0) == 0


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(repo)r)
from qgs_b200.ensemble import shard_bounds, combine_moments, gather_states, spectrum_from_local_exponents
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
states = np.random.default_rng(0).standard_normal((1001, 36))
lo, hi = shard_bounds(len(states), 2, rank)
mine = states[lo:hi]
mean, var, count = combine_moments(mine.sum(axis=0), (mine ** 2).sum(axis=0), len(mine))
assert count == 1001
assert np.allclose(mean, states.mean(axis=0), rtol=1e-12, atol=1e-14)
assert np.allclose(var, states.var(axis=0), rtol=1e-10)
full = gather_states(mine)
assert np.array_equal(full, states)
# Lyapunov spectrum of a sharded ensemble: member/time means combined over the ranks
exps = np.random.default_rng(1).standard_normal((1001, 5, 7)) + np.arange(5)[None, :, None]
spec, sem = spectrum_from_local_exponents(exps[lo:hi], 1001)
per_member = exps.mean(axis=2)
assert np.allclose(spec, per_member.mean(axis=0), rtol=1e-12)
assert np.allclose(sem, per_member.std(axis=0) / np.sqrt(1000.), rtol=1e-9)
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
"""


def test_sharded_moments_and_gather_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    port = 29600 + os.getpid() % 200
    script.write_text(_GLOO_WORKER % {"repo": REPO, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_bench_reference_arm_contract():
    """--impl reference prints one JSON line with the contract's keys (bounded CPU sample)."""
    import json
    env = dict(os.environ, QGSB_BENCH_CPU_SECONDS="0.5", QGSB_BENCH_CPU="port")     # the port leg: fast and always there
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "member-steps/s" and line["value"] > 1e4
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_overlay_serves_hot_path_modules_and_leaves_the_rest_to_the_reference():
    ref = os.environ.get("QGS_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "qgs")):
        pytest.skip("reference checkout not available (GPU box)")
    code = (
        "import warnings; warnings.filterwarnings('ignore')\n"
        "import qgs\n"
        "from qgs.params.params import QgParams\n"
        "from qgs.integrators.integrator import RungeKuttaIntegrator, RungeKuttaTglsIntegrator\n"
        "from qgs.functions.tendencies import create_tendencies\n"
        "from qgs.toolbox.lyapunov import LyapunovsEstimator, CovariantLyapunovsEstimator\n"
        "from qgs.integrators.integrate import integrate_runge_kutta\n"
        "from qgs.functions.sparse_mul import sparse_mul3\n"
        "import qgs.functions.util as u, qgs.integrators.statistics as st\n"
        "assert RungeKuttaIntegrator.__module__ == 'qgs_b200.integrators.integrator'\n"
        "assert create_tendencies.__module__ == 'qgs_b200.functions.tendencies'\n"
        "assert LyapunovsEstimator.__module__ == 'qgs_b200.toolbox.lyapunov'\n"
        "assert st.TrajectoriesStatistics.__module__ == 'qgs_b200.integrators.statistics'\n"
        "assert QgParams.__module__ == 'qgs.params.params' and %r in u.__file__\n"
        "print('overlay ok')\n" % (ref,))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REPO, "overlay"), ref]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "overlay ok" in out.stdout, out.stderr[-2000:]


def test_reference_arm_runs_the_installed_reference_and_agrees_with_the_oracle():
    """bench.py --impl reference drives the UNMODIFIED reference from baseline/_ref (numba f + RK4 + worker pool);
    the same object integrates a few members here and must agree with the C oracle -- a live pin of the oracle."""
    import importlib.util
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "qgs")) or importlib.util.find_spec("numba") is None:
        pytest.skip("baseline/_ref is not installed (python -m pip install --no-index --no-build-isolation --no-deps "
                    "--target baseline/_ref /root/reference)")
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "os.cpu_count = lambda: 2\n"
        "import bench, oracle\n"
        "ref = bench.ReferenceNumba()\n"
        "ic = bench.initial_conditions(0, 6)\n"
        "ref.integrator.integrate(0., 5., 0.1, ic=ic, write_steps=5)\n"
        "t, traj = ref.integrator.get_trajectories()\n"
        "ref.close()\n"
        "T = oracle.Tensor.from_npz(bench.TENSOR)\n"
        "b, c, a = oracle.rk4_tableau()\n"
        "tv = np.concatenate((np.arange(0., 5., 0.1), [5.]))\n"
        "o = oracle.integrate_runge_kutta_jit(T, tv, ic, 1, 5, b, c, a)\n"
        "err = np.max(np.abs(traj - o)) / np.max(np.abs(o))\n"
        "assert traj.shape == o.shape and err < 1e-13, err\n"
        "print('reference arm ok', err)\n" % REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0 and "reference arm ok" in out.stdout, out.stderr[-3000:]


def test_resident_ensemble_record_times_follow_the_integrator_rules():
    """DeviceEnsemble's record-time vector (streamed trajectories, record moments) == the time vector
    RungeKuttaIntegrator.get_trajectories returns (integrator.py:409-424), forward, backward, ragged, write_steps 0."""
    from qgs_b200.ensemble import DeviceEnsemble
    from qgs_b200.integrators.integrate import returned_time
    for t_end in (1.0, 1.35, 2.05):
        time = np.concatenate((np.arange(0., t_end, 0.1), [t_end]))
        for ws in (0, 1, 2, 3, 7, 50):
            for forward in (True, False):
                ref = np.atleast_1d(np.asarray(returned_time(time, 0., t_end, forward, ws)))
                got = DeviceEnsemble._record_times(time, ws, forward)
                assert np.array_equal(ref, got), (t_end, ws, forward)


# ---- tensor hand-off file and parameter fingerprint ----------------------------------------
def test_tensor_file_round_trip_and_validation(tmp_path):
    from qgs_b200.functions import tensor_cache
    z = np.load(os.path.join(GOLDEN, "tensor_T4.npz"))
    path = tensor_cache.save_tensor(str(tmp_path / "t4.npz"), int(z["ndim"]), z["coo"].astype(np.int64), z["val"],
                                    z["jcoo"], z["jval"])
    ndim, coo, val, jcoo, jval = tensor_cache.load_tensor(path)
    assert ndim == int(z["ndim"]) and coo.dtype == np.int32 and coo.shape[1] == 5
    assert np.array_equal(coo, z["coo"]) and np.array_equal(val, z["val"])
    assert np.array_equal(jcoo, z["jcoo"]) and np.array_equal(jval, z["jval"])
    assert os.listdir(tmp_path) == ["t4.npz"]                       # no temporary file left behind
    # a tensor without Jacobian part (create_atmo_thermo_tendencies) round-trips too
    thermo = tensor_cache.save_tensor(str(tmp_path / "thermo.npz"), ndim, coo, val, np.zeros((0, 5), np.int32), np.zeros(0))
    assert tensor_cache.load_tensor(thermo)[3].shape == (0, 5) and len(tensor_cache.load_tensor(thermo)[4]) == 0
    os.remove(thermo)
    # the golden fixtures are in the same format
    assert tensor_cache.load_tensor(os.path.join(GOLDEN, "tensor_maooam36.npz"))[0] == 36
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, ndim=3, coo=np.zeros((4, 3), np.int32), val=np.zeros(3), jcoo=np.zeros((0, 3), np.int32), jval=np.zeros(0))
    with pytest.raises(ValueError, match="not a tendencies tensor"):
        tensor_cache.load_tensor(bad)
    np.savez(bad, ndim=3, coo=np.full((4, 3), 9, np.int32), val=np.zeros(4), jcoo=np.zeros((0, 3), np.int32), jval=np.zeros(0))
    with pytest.raises(ValueError, match="outside"):
        tensor_cache.load_tensor(bad)


def test_parameter_fingerprint_tracks_values_not_identities(tmp_path, monkeypatch):
    from qgs_b200.functions import tensor_cache

    class Scaled(float):                                            # like qgs Parameter: a float with attributes
        def __new__(cls, value, units=""):
            obj = float.__new__(cls, value)
            obj.units = units
            return obj

    class Block(object):
        def __init__(self, kd, modes):
            self.kd = Scaled(kd, "[1/s]")
            self.modes = np.array(modes)
            self.nested = {"eps": 0.7, "labels": ["a", "b"], "fn": np.sin}
            self.me = self                                            # cycle

    a, b = Block(0.029, [[1, 1], [1, 2]]), Block(0.029, [[1, 1], [1, 2]])
    assert tensor_cache.fingerprint(a) == tensor_cache.fingerprint(b)      # equal values, different objects
    assert tensor_cache.fingerprint(a) != tensor_cache.fingerprint(a, "thermo")
    for change in (lambda o: setattr(o, "kd", Scaled(0.03, "[1/s]")), lambda o: setattr(o, "kd", Scaled(0.029, "[1/d]")),
                   lambda o: o.modes.__setitem__((1, 1), 3), lambda o: o.nested.__setitem__("eps", 0.71),
                   lambda o: o.nested["labels"].append("c"), lambda o: o.nested.__setitem__("fn", np.cos)):
        c = Block(0.029, [[1, 1], [1, 2]])
        change(c)
        assert tensor_cache.fingerprint(c) != tensor_cache.fingerprint(a)
    monkeypatch.delenv("QGSB_TENSOR_CACHE", raising=False)
    assert tensor_cache.cache_file(a) is None
    monkeypatch.setenv("QGSB_TENSOR_CACHE", str(tmp_path / "cache"))
    path = tensor_cache.cache_file(a)
    assert path == tensor_cache.cache_file(b) and os.path.isdir(os.path.dirname(path))


def test_fingerprint_of_reference_parameter_objects():
    """On the reference's own QgParams (from the installed baseline/_ref): equal set-ups built separately share a cache
    entry, every changed scalar / mode set / basis kind gets its own."""
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "qgs")):
        pytest.skip("baseline/_ref is not installed")
    code = (
        "import warnings; warnings.filterwarnings('ignore')\n"
        "from qgs_b200 import compat; compat.install()\n"
        "from qgs.params.params import QgParams\n"
        "from qgs_b200.functions.tensor_cache import fingerprint\n"
        "def make(kd=0.029, ny=2, mode='analytic', insolation=103.3333):\n"
        "    p = QgParams()\n"
        "    p.set_atmospheric_channel_fourier_modes(2, ny, mode=mode)\n"
        "    p.set_oceanic_basin_fourier_modes(2, 4, mode=mode)\n"
        "    p.set_params({'kd': kd, 'kdp': 0.0290, 'n': 1.5, 'r': 1.e-7, 'h': 136.5, 'd': 1.1e-7})\n"
        "    p.atemperature_params.set_params({'eps': 0.7, 'T0': 289.3, 'hlambda': 15.06})\n"
        "    p.gotemperature_params.set_params({'gamma': 5.6e8, 'T0': 301.46})\n"
        "    p.atemperature_params.set_insolation(insolation, 0)\n"
        "    p.gotemperature_params.set_insolation(310., 0)\n"
        "    return p\n"
        "base = fingerprint(make())\n"
        "assert base == fingerprint(make())\n"
        "others = [fingerprint(make(kd=0.03)), fingerprint(make(ny=3)), fingerprint(make(mode='symbolic')),\n"
        "          fingerprint(make(insolation=104.))]\n"
        "assert len(set(others + [base])) == 5\n"
        "assert fingerprint(make(mode='symbolic')) == others[2]\n"
        "print('fingerprints ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REPO, ref_dir]))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "fingerprints ok" in res.stdout, res.stderr[-3000:]


def test_overlay_adapts_the_vertical_velocity_diagnostic_on_import(tmp_path):
    """qgs/diagnostics/wind.py:705-714 hands the tendencies to @njit code.  Under the overlay the module is the
    reference's file, loaded by the reference's loader, with `_compute_omega_term` replaced by the batched evaluation of
    device tendencies (anything else is rejected: no CPU path)."""
    from conftest import write_plot_stubs
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "qgs")):
        pytest.skip("baseline/_ref is not installed")
    code = (
        "import warnings; warnings.filterwarnings('ignore')\n"
        "import sys, numpy as np\n"
        "from numba import njit\n"
        "import qgs.diagnostics.wind as wind\n"
        "from qgs_b200 import overlay_hooks\n"
        "assert wind._qgsb_patched and not hasattr(wind._compute_omega_term, 'py_func')\n"
        "assert wind.__file__.endswith('wind.py') and 'baseline' in wind.__file__\n"
        "assert wind.__loader__.get_filename() == wind.__file__\n"
        "@njit\n"
        "def f(t, x): return 2. * x + t\n"
        "@njit\n"
        "def g(t, x): return x * x\n"
        "time = np.arange(5) * 0.1\n"
        "data = np.random.default_rng(0).random((7, 5))\n"
        "try:\n"
        "    wind._compute_omega_term(time, data, f, g)\n"
        "    raise SystemExit('numba functions accepted: that would be a CPU path')\n"
        "except TypeError as exc:\n"
        "    assert 'no CPU fallback' in str(exc)\n"
        "out = overlay_hooks.omega_term(time, data, lambda t, x: 2. * x, lambda t, x: x * x)\n"
        "assert out.shape == (7, 5) and out.flags.c_contiguous and np.allclose(out, 2. * data - data * data)\n"
        "try:\n"
        "    overlay_hooks.omega_term(time, data[:, :4], f, g)\n"
        "    raise SystemExit('record count mismatch accepted')\n"
        "except ValueError:\n"
        "    pass\n"
        "overlay_hooks.install(); overlay_hooks.install()\n"
        "assert sum(isinstance(m, overlay_hooks.PatchAfterImport) for m in sys.meta_path) == 1\n"
        "print('hook ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([write_plot_stubs(tmp_path / "stubs"),
                                                       os.path.join(REPO, "overlay"), ref_dir]))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "hook ok" in res.stdout, res.stderr[-3000:]


# ---- generated modules: versioned, and safe to build from several processes at once -------------------------
def test_plugin_built_against_another_kernel_table_layout_is_refused(tmp_path):
    """A module left in the JIT cache by an older build must not be read with the new SpecKernels layout:
    qgsb_load_plugin checks the layout version the module was compiled with (no device needed)."""
    import subprocess
    from qgs_b200 import _lib
    src = tmp_path / "stale.cu"
    src.write_text('#include "spec_registry.h"\n'
                   'static const qgsb::SpecKernels k = {QGSB_SPEC_ABI + 1u, 1ULL, 3, 3, 1, "stale", nullptr, nullptr, '
                   'nullptr, 0, nullptr, 0ULL};\n'
                   'extern "C" __attribute__((visibility("default"))) const qgsb::SpecKernels *qgsb_plugin_kernels(void) '
                   '{ return &k; }\n')
    so = tmp_path / "stale.so"
    subprocess.check_call(["nvcc", "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-I", os.path.join(REPO, "qgs_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = _lib.load()
    assert lib.qgsb_load_plugin(str(so).encode()) != 0
    msg = lib.qgsb_last_error().decode()
    assert "layout" in msg and "rebuild" in msg


def test_concurrent_plugin_builds_of_one_tensor_do_not_collide(tmp_path):
    """Under torchrun every rank misses the JIT cache at once: builders write private scratch files, queue on
    a lock and the module is compiled once; every process ends up with the same complete shared object."""
    import subprocess
    import sys
    worker = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
              "from qgs_b200 import codegen\n"
              "coo = np.array([[1, 0, 1], [1, 0, 2], [2, 0, 1], [2, 0, 2], [2, 1, 3], [3, 0, 3], [3, 1, 2]])\n"
              "val = np.array([-10., 10., 28., -1., -1., -2.5, 1.])\n"
              "print(codegen.build_plugin(3, 3, coo, val, part='rk'))\n" % REPO)
    env = dict(os.environ, QGSB_JIT_DIR=str(tmp_path))
    procs = [subprocess.Popen([sys.executable, "-c", worker], env=env, stdout=subprocess.PIPE, text=True)
             for _ in range(3)]
    paths = {p.communicate()[0].strip() for p in procs}
    assert all(p.returncode == 0 for p in procs)
    assert len(paths) == 1 and paths != {"None"}
    path = paths.pop()
    from qgs_b200 import codegen
    assert os.path.exists(path) and codegen.source_key() in os.path.basename(path)
    left = sorted(os.listdir(str(tmp_path)))
    assert [x for x in left if x.endswith(".so")] == [os.path.basename(path)]
    assert not [x for x in left if x.endswith(".tmp") or x.endswith(".cu")]        # scratch files are gone
    from qgs_b200 import _lib
    assert _lib.load().qgsb_load_plugin(path.encode()) == 0


# ---- coordinate-based `sparse` stand-in ----------------------------------------------------------
def _sparse():
    import importlib.util
    spec = importlib.util.spec_from_file_location("qgsb_sparse_standin",
                                                  os.path.join(REPO, "qgs_b200", "compat", "sparse", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_coo_standin_matches_dense_numpy_on_every_operation_qgs_uses():
    """COO / DOK keep coordinate lists (memory ~ nnz), and every operation the reference's tensor construction performs on
    them (qgtensor.py:188-272, 657-746, 969-1005; analytic.py:131-216) gives what numpy gives on the dense arrays."""
    sp = _sparse()
    rng = np.random.default_rng(0)

    def rand(shape, density=0.3):
        a = rng.standard_normal(shape)
        a[rng.random(shape) > density] = 0.
        return a

    a3, m2, v1, w1 = rand((5, 6, 7)), rand((6, 6)), rand((6,)), rand((7,), 0.9)
    A3, M2, V1, W1 = sp.COO(a3), sp.COO(m2), sp.COO(v1), sp.COO(w1)
    # entries in lexicographic order, zeros not stored
    assert np.array_equal(A3.coords, np.array(np.nonzero(a3))) and np.array_equal(A3.data, a3[np.nonzero(a3)])
    assert A3.nnz == np.count_nonzero(a3) and A3.shape == a3.shape and A3.ndim == 3
    # slicing: integers and unit-step slices, partial keys, scalars
    for key in ((2,), (slice(1, None), 3), (slice(1, None), 3, slice(2, None)), (slice(None), 2, 4), (4, 5, 6),
                (slice(None), slice(None), 0), (-1, slice(None), slice(None))):
        got = A3[key]
        want = a3[key]
        if np.ndim(want) == 0:
            assert isinstance(got, float) and got == float(want)
        else:
            assert got.shape == want.shape and np.array_equal(got.todense(), want)
    # products: vector @ matrix @ vector -> scalar, matrix @ vector, vector @ sliced tensor
    assert np.isclose(V1 @ M2 @ sp.COO(v1), v1 @ m2 @ v1) and isinstance(V1 @ M2 @ V1, float)
    assert np.array_equal((M2 @ V1).todense(), m2 @ v1) and np.array_equal((v1 @ M2).todense(), v1 @ m2)
    assert np.allclose((V1 @ A3[1, :, :]).todense(), v1 @ a3[1])
    assert np.isclose(M2[2, :] @ A3[2, :, 3], m2[2] @ a3[2, :, 3])
    # tensordot(axes=1): last axis against the first, also through the coordinate join used above the dense limit
    t5 = rand((6, 3, 4, 3, 2), 0.2)
    want = np.tensordot(v1, t5, axes=1)
    assert np.allclose(sp.tensordot(V1, sp.COO(t5), axes=1).todense(), want)
    old = sp.DENSE_LIMIT
    try:
        sp.DENSE_LIMIT = 0
        assert np.allclose(sp.tensordot(sp.COO(v1), sp.COO(t5), axes=1).todense(), want)
        assert np.allclose((sp.COO(m2) @ sp.COO(v1)).todense(), m2 @ v1)
        assert np.isclose(sp.COO(v1) @ sp.COO(m2) @ sp.COO(v1), v1 @ m2 @ v1)
        assert np.array_equal(sp.COO(a3)[slice(1, None), 3, slice(2, None)].todense(), a3[1:, 3, 2:])
        assert sp.COO(a3)[4, 5, 6] == a3[4, 5, 6]
    finally:
        sp.DENSE_LIMIT = old
    # arithmetic, swapaxes, duplicates summed in input order, pruning
    assert np.array_equal((A3 + A3.swapaxes(1, 2).swapaxes(1, 2)).todense(), 2 * a3)
    assert np.array_equal(A3.swapaxes(0, 2).todense(), a3.swapaxes(0, 2))
    assert np.array_equal((2.5 * A3).todense(), 2.5 * a3) and np.array_equal((-A3).todense(), -a3)
    assert np.array_equal((A3 * 0.5).todense(), a3 * 0.5) and np.array_equal((A3 - A3).coords.shape, (3, 0))
    c = sp.COO(np.array([[1, 1, 0, 1], [2, 2, 0, 2]]), np.array([1., 1e-17, 3., -1.]), shape=(3, 3), prune=True)
    assert np.array_equal(c.coords, [[0, 1], [0, 2]]) and np.array_equal(c.data, [3., (1. + 1e-17) - 1.]) or c.nnz == 1
    z = sp.zeros((4, 4, 4), format='coo')
    assert z.nnz == 0 and z.coords.shape == (3, 0) and z.coords.size == 0
    # DOK: item get / set / in-place updates with integer tuples, conversion
    d = sp.zeros((3, 4), dtype=float, format='dok')
    d[1, 2] = 2.
    d[1, 2] -= 0.5
    d[(0, 0)] += V1 @ M2 @ V1
    d[2, 3] = 0.
    assert d[1, 2] == 1.5 and d[2, 2] == 0.0 and d.nnz == 2
    dc = d.to_coo()
    assert isinstance(dc, sp.COO) and dc.shape == (3, 4) and np.array_equal(dc.coords, [[0, 1], [0, 2]])
    assert dc.to_coo() is dc


def test_coo_standin_holds_a_large_basis_rank5_tensor_in_kilobytes():
    """The point of coordinate storage: a rank-5 tensor over 229 indices (the 6x6 model with T^4 terms) is 229**5
    doubles = 5 TB dense; as a coordinate list it is proportional to its entries, and the operations of
    qgtensor.py:657-746 (gathering, Jacobian by swapped axes, upper-triangular simplification) run on the lists."""
    sp = _sparse()
    rng = np.random.default_rng(1)
    n1, nnz = 229, 20000
    coords = np.vstack((rng.integers(1, n1, (1, nnz)), np.sort(rng.integers(0, n1, (4, nnz)), axis=0)))
    t = sp.COO(coords, rng.standard_normal(nnz), shape=(n1,) * 5)
    assert t.nnz <= nnz and t.coords.nbytes + t.data.nbytes < 1 << 20
    jac = t.copy()
    for i in range(1, 4):
        jac += t.swapaxes(1, i + 1)
    assert jac.shape == t.shape and jac.nnz >= t.nnz
    cs = jac.coords.copy()
    cs[1:, :] = np.sort(cs[1:, :], axis=0)
    upp = sp.COO(cs, jac.data.copy(), shape=jac.shape, prune=True)
    # a term with four distinct factors appears under each of them in the Jacobian: folding back gives 4x the tensor
    same = {tuple(c): v for c, v in zip(t.coords.T.tolist(), t.data.tolist())}
    got = {tuple(c): v for c, v in zip(upp.coords.T.tolist(), upp.data.tolist())}
    assert set(got) == set(same)
    assert all(abs(got[k] - 4. * same[k]) <= 1e-12 * abs(same[k]) for k in same)


@pytest.mark.skipif(not os.path.isdir("/root/reference/qgs"), reason="needs the reference checkout (build container)")
def test_reference_tensor_construction_on_the_coo_standin_reproduces_the_fixtures():
    """create_tendencies of the UNMODIFIED reference on top of the stand-in rebuilds the MAOOAM-36 and RP-20 fixtures bit
    for bit (the other five configurations: tests/golden/check_tensors.py, profiles/r02_coo_rebuild.log)."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "tests", "golden", "check_tensors.py"), "maooam36", "rp"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if "identical=" in ln]
    assert len(lines) == 2 and all("identical=True" in ln for ln in lines), out.stdout
