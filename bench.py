#!/usr/bin/env python
"""bench.py -- ensemble RK4 member-steps/s, MAOOAM-36 (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One bench "step" = one pass of the hot path over one batch: every rank integrates its shard of
MEMBERS_PER_GPU MAOOAM-36 members (qgs_maooam.py parameters, tests/golden/tensor_maooam36.npz) over
STEPS_PER_LAUNCH classic-RK4 time steps of dt = 0.1 with write_steps = 0 -- one fused kernel launch.

  value     device-timed (CUDA events on the launching stream) whole-job member-steps/s with the
            ensemble already resident in HBM; max over ranks of the summed launch times.
  e2e       the same work through the reference-facing call (qgsb_rk_integrate == what
            RungeKuttaIntegrator.integrate()/get_trajectories() run) with PINNED HOST buffers: the
            host->device copy of the initial conditions and the device->host copy of the final states
            are inside the timed region, every step.
  roofline  FP64: algorithmic flops (4132 per member-step) / launch time against the
            DFMA peak measured on this device in this run (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  the UNMODIFIED reference (numba + its multiprocessing pool, all host cores) from the git-ignored
            baseline/_ref install on a bounded sample; the C oracle port (oracle/qgs_oracle.c, all host threads) is
            reported next to it as cpu_port, and replaces it (kind "port") when baseline/_ref or numba is missing.

  lyapunov  BASELINE.json's fifth configuration as extra keys of the same line: the Benettin loop (tangent-linear RK4
            of the 36 x 36 basis + Householder QR per step) of LyapunovsEstimator.compute_lyapunovs(0, 100, 200, 0.1,
            0.1, write_steps=10) on 8192 members per GPU -- device-timed member-steps/s of qgsb_lyap_benettin with
            its own FP64 roofline (525 532 flops per member-step), e2e through the Python class
            (host arrays in, trajectory + exponents out), and the reference's LyapunovsEstimator on all host cores.
  strong    the SAME total ensemble (2^20 members) split over the N GPUs: whole-job member-steps/s, device-timed.
  suite     the other configurations of BASELINE.json (RP-20, dynamic-T, T4, 6x6 large basis) on ONE GPU (rank 0):
            device-timed RK4 member-steps/s and the FP64 roofline fraction with each tensor's own algorithmic flops
            (SURVEY.md section 8d).  Parity of these configurations is what tests/test_gpu_parity.py checks.

--impl reference times the reference's own CPU implementation of the path: RungeKuttaIntegrator from
baseline/_ref through its public API (else the oracle port).  Under torchrun rank 0 alone runs it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "ensemble RK4 member-steps/sec, MAOOAM-36"
UNIT = "member-steps/s"
MEMBERS_PER_GPU = 1 << 20
STEPS_PER_LAUNCH = 1000
DT = 0.1
NDIM = 36
FLOPS_PER_MEMBER_STEP = 4132          # 4 * sum_nnz(p + 1) + 14 * ndim, SURVEY.md section 8d
TENSOR = os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz")
# the Lyapunov configuration (BASELINE.json configs[4])
LYAP_MEMBERS_PER_GPU = 8192
LYAP_ARGS = (0., 100., 200., 0.1, 0.1)      # t0, tw, t, dt, mdt of compute_lyapunovs
LYAP_WRITE_STEPS = 10
LYAP_NVEC = 36
LYAP_STEPS = 2000                            # Benettin steps per member: 1000 of convergence + 1000 recorded
# 4 (F_f + 2 jnnz + 2 n^2 m) + 14 n + 14 n m + 4 n m^2 - 4/3 m^3 with F_f = 907, jnnz = 699, n = m = 36
LYAP_FLOPS_PER_MEMBER_STEP = 525532


def workload_config(n_gpus):
    return {"workload": "MAOOAM-36 (qgs_maooam.py parameters, nnz 351), %d members/GPU x %d classic-RK4 steps per "
                        "launch, dt=0.1, write_steps=0" % (MEMBERS_PER_GPU, STEPS_PER_LAUNCH),
            "members_per_gpu": MEMBERS_PER_GPU, "steps_per_launch": STEPS_PER_LAUNCH, "n_dim": NDIM,
            "parallelism": "members sharded over %d GPU(s), no inter-GPU traffic during integration" % n_gpus,
            "l2": "ensemble state 302 MB per GPU > 126 MB L2 (inputs larger than L2)"}


def initial_conditions(rank, members):
    rng = np.random.default_rng(21217 + rank)
    return rng.random((members, NDIM)) * 0.01


def time_vector(steps):
    return np.concatenate((np.arange(0., steps * DT, DT), np.full((1,), steps * DT)))


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range seen
        thr = 0.5 * (min(power) + max(power))
        loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_rate(target_seconds=12.0):
    target_seconds = float(os.environ.get("QGSB_BENCH_CPU_SECONDS", target_seconds))
    """member-steps/s of the CPU port with all host threads on a bounded sample of the same workload."""
    import oracle
    T = oracle.Tensor.from_npz(TENSOR)
    b, c, a = oracle.rk4_tableau()
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    time_v = time_vector(STEPS_PER_LAUNCH)
    probe = initial_conditions(0, 512)
    t0 = time.perf_counter()
    oracle.integrate_runge_kutta_jit(T, time_v, probe, 1, 0, b, c, a)       # warm-up + calibration
    est = 512 * STEPS_PER_LAUNCH / (time.perf_counter() - t0)
    members = int(min(max(est * target_seconds / STEPS_PER_LAUNCH // 1024 * 1024, 1024), 1 << 18))
    ic = initial_conditions(0, members)
    t0 = time.perf_counter()
    oracle.integrate_runge_kutta_jit(T, time_v, ic, 1, 0, b, c, a)
    wall = time.perf_counter() - t0
    return {"value": members * STEPS_PER_LAUNCH / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d members x %d RK4 steps (write_steps=0), oracle/qgs_oracle.c with %d pthreads, %.1f s wall"
                      % (members, STEPS_PER_LAUNCH, cores, wall)}, members, wall


REF_DIR = os.path.join(REPO, "baseline", "_ref")


class ReferenceNumba(object):
    """The UNMODIFIED reference (pip-installed from /root/reference into the git-ignored baseline/_ref, see
    DESIGN.md section 6) on its own stock path: QgParams of qgs_maooam.py:78-92 -> create_tendencies ->
    RungeKuttaIntegrator(num_threads = host cores).integrate(..., write_steps=0) -> get_trajectories().
    pydata `sparse` / `pebble`, which the reference imports while it BUILDS the tensor, are absent from the image;
    qgs_b200.compat provides stand-ins for that setup step only -- f, the RK4 loop and the worker pool are the
    reference's numba / multiprocessing code."""

    def __init__(self):
        if not os.path.isdir(os.path.join(REF_DIR, "qgs")):
            raise RuntimeError("baseline/_ref/qgs is not installed")
        import warnings
        warnings.filterwarnings("ignore")
        from qgs_b200 import compat
        compat.install()
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import numba  # noqa: F401  (fail early when the reference's JIT is missing)
        from qgs.params.params import QgParams
        from qgs.functions.tendencies import create_tendencies
        from qgs.integrators.integrator import RungeKuttaIntegrator
        import qgs
        if os.path.abspath(os.path.dirname(qgs.__file__)) != os.path.abspath(os.path.join(REF_DIR, "qgs")):
            raise RuntimeError("qgs resolved to %s, not to baseline/_ref" % qgs.__file__)
        p = QgParams()
        p.set_atmospheric_channel_fourier_modes(2, 2)
        p.set_oceanic_basin_fourier_modes(2, 4)
        p.set_params({'kd': 0.0290, 'kdp': 0.0290, 'n': 1.5, 'r': 1.e-7, 'h': 136.5, 'd': 1.1e-7})
        p.atemperature_params.set_params({'eps': 0.7, 'T0': 289.3, 'hlambda': 15.06, })
        p.gotemperature_params.set_params({'gamma': 5.6e8, 'T0': 301.46})
        p.atemperature_params.set_insolation(103.3333, 0)
        p.gotemperature_params.set_insolation(310., 0)
        f, Df = create_tendencies(p)
        self.f, self.Df = f, Df
        self.cores = os.cpu_count() or 1
        self.integrator = RungeKuttaIntegrator(num_threads=self.cores)
        self.integrator.set_func(f)
        # every worker JIT-compiles the integrator on its first trajectory
        self.run(4 * self.cores, 10)

    def run(self, members, steps):
        ic = initial_conditions(0, members)
        t0 = time.perf_counter()
        self.integrator.integrate(0., steps * DT, DT, ic=ic, write_steps=0)
        self.integrator.get_trajectories()
        return time.perf_counter() - t0

    def rate(self, target_seconds):
        wall = self.run(8 * self.cores, STEPS_PER_LAUNCH)
        est = 8 * self.cores * STEPS_PER_LAUNCH / wall
        members = int(min(max(est * target_seconds / STEPS_PER_LAUNCH // self.cores * self.cores, self.cores), 1 << 16))
        wall = self.run(members, STEPS_PER_LAUNCH)
        return {"value": members * STEPS_PER_LAUNCH / wall, "unit": UNIT, "cores": self.cores, "kind": "reference",
                "sample": "%d members x %d RK4 steps (write_steps=0), unmodified qgs RungeKuttaIntegrator (numba + %d "
                          "worker processes) from baseline/_ref, %.1f s wall" % (members, STEPS_PER_LAUNCH, self.cores,
                                                                               wall)}, members, wall

    def close(self):
        try:
            self.integrator.terminate()
        except Exception:
            pass


class ReferenceLyapunov(object):
    """The UNMODIFIED reference's LyapunovsEstimator (qgs/toolbox/lyapunov.py:41-468: numba + one worker process per
    host core) from baseline/_ref on the Lyapunov configuration: compute_lyapunovs(0, 100, 200, 0.1, 0.1,
    write_steps=10), 36 vectors."""

    def __init__(self, ref):
        from qgs.toolbox.lyapunov import LyapunovsEstimator
        self.cores = ref.cores
        self.est = LyapunovsEstimator(num_threads=self.cores)
        self.est.set_func(ref.f, ref.Df)
        # every worker JIT-compiles the Benettin loop on its first member
        self.est.compute_lyapunovs(0., 0.2, 0.4, 0.1, 0.1, initial_conditions(0, self.cores), write_steps=1)
        self.est.get_lyapunovs()

    def rate(self, members=None):
        members = members or 16 * self.cores
        ic = initial_conditions(0, members)
        t0 = time.perf_counter()
        self.est.compute_lyapunovs(*LYAP_ARGS, ic, write_steps=LYAP_WRITE_STEPS)
        self.est.get_lyapunovs()
        wall = time.perf_counter() - t0
        return {"value": members * LYAP_STEPS / wall, "unit": UNIT, "cores": self.cores, "kind": "reference",
                "sample": "%d members x %d Benettin steps, 36 vectors, unmodified qgs LyapunovsEstimator (numba + %d "
                          "worker processes) from baseline/_ref, %.1f s wall" % (members, LYAP_STEPS, self.cores, wall)}

    def close(self):
        try:
            self.est.terminate()
        except Exception:
            pass


def lyapunov_port_rate(target_seconds=8.0):
    """The C oracle port of the Benettin loop on all host threads (used when the reference cannot run)."""
    import oracle
    T = oracle.Tensor.from_npz(TENSOR)
    b, c, a = oracle.rk4_tableau()
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    members = cores
    ic = initial_conditions(0, members)
    rng = np.random.default_rng(7)
    q0 = np.stack([np.linalg.qr(rng.random((NDIM, LYAP_NVEC)))[0] for _ in range(members)])
    pre = np.concatenate((np.arange(LYAP_ARGS[0], LYAP_ARGS[1], LYAP_ARGS[3]), [LYAP_ARGS[1]]))
    tim = np.concatenate((np.arange(LYAP_ARGS[1], LYAP_ARGS[2], LYAP_ARGS[3]), [LYAP_ARGS[2]]))
    t0 = time.perf_counter()
    oracle.compute_backward_lyap(T, pre, tim, LYAP_ARGS[4], ic, LYAP_NVEC, LYAP_WRITE_STEPS, False, 1., b, c, a, q0, None)
    wall = time.perf_counter() - t0
    return {"value": members * LYAP_STEPS / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d members x %d Benettin steps, 36 vectors, oracle/qgs_oracle.c with %d pthreads, %.1f s wall"
                      % (members, LYAP_STEPS, cores, wall)}


def cpu_legs(rk_seconds, with_port=True):
    """All host-side baselines of one bench line, run BEFORE the GPU work and before any process group exists (the
    other ranks must not spin in a collective while rank 0 times CPU code): the reference's RungeKuttaIntegrator and
    LyapunovsEstimator from baseline/_ref when they can run here, else the C port; plus the C port of the RK4 path
    as a second data point."""
    base = port = lyap = None
    ref = None
    if os.environ.get("QGSB_BENCH_CPU", "") != "port":
        try:
            ref = ReferenceNumba()
            base, _, _ = ref.rate(rk_seconds)
        except Exception as exc:  # missing install / numba: say so and use the port
            sys.stderr.write("reference numba path unavailable (%s); timing the C port instead\n" % (exc,))
            base = None
        if ref is not None and base is not None:
            rl = None
            try:
                rl = ReferenceLyapunov(ref)
                lyap = rl.rate()
            except Exception as exc:
                sys.stderr.write("reference LyapunovsEstimator unavailable (%s); timing the C port instead\n" % (exc,))
            finally:
                if rl is not None:
                    rl.close()
        if ref is not None:
            ref.close()
    if base is None:
        base, _, _ = cpu_rate(rk_seconds)
    elif with_port:
        port, _, _ = cpu_rate(6.0)      # the C oracle port as a second data point (it is faster than numba)
    if lyap is None:
        lyap = lyapunov_port_rate()
    return base, port, lyap


def reference_or_port(target_seconds):
    """(baseline dict, members, wall): the reference's numba path when baseline/_ref can run here, else the C port."""
    if os.environ.get("QGSB_BENCH_CPU", "") != "port":
        ref = None
        try:
            ref = ReferenceNumba()
            return ref.rate(target_seconds)
        except Exception as exc:  # missing install / numba: say so and use the port
            sys.stderr.write("reference numba path unavailable (%s); timing the C port instead\n" % (exc,))
        finally:
            if ref is not None:
                ref.close()
    return cpu_rate(target_seconds)


def run_reference(args, rank):
    if rank != 0:
        return
    rates, walls = [], []
    base = None
    ref = None
    lyap = None
    if os.environ.get("QGSB_BENCH_CPU", "") != "port":
        try:
            ref = ReferenceNumba()
        except Exception as exc:
            sys.stderr.write("reference numba path unavailable (%s); timing the C port instead\n" % (exc,))
    for i in range(args.warmup + args.steps):
        base, members, wall = ref.rate(6.0) if ref is not None else cpu_rate(target_seconds=6.0)
        if i >= args.warmup:
            rates.append(base["value"])
            walls.append(wall)
    if ref is not None:
        rl = None
        try:
            rl = ReferenceLyapunov(ref)
            lyap = rl.rate()
        except Exception as exc:
            sys.stderr.write("reference LyapunovsEstimator unavailable (%s); timing the C port instead\n" % (exc,))
        finally:
            if rl is not None:
                rl.close()
        ref.close()
    if lyap is None:
        lyap = lyapunov_port_rate()
    value = float(np.mean(rates))
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "lyapunov": {"metric": LYAP_METRIC, "value": lyap["value"], "unit": UNIT, "cpu_baseline": lyap,
                         "e2e": {"value": lyap["value"], "unit": UNIT}},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
LYAP_METRIC = "Benettin member-steps/sec (tangent-linear RK4 of 36 vectors + QR per step), MAOOAM-36"


def lyapunov_leg(lib, _lib, f, Df, rank, world, barrier, repeats):
    """BASELINE.json configs[4]: LyapunovsEstimator.compute_lyapunovs(0, 100, 200, 0.1, 0.1, write_steps=10) on
    LYAP_MEMBERS_PER_GPU members per rank, 36 vectors.  Returns (device seconds, e2e seconds, launches, finite) of
    `repeats` calls on this rank."""
    import ctypes
    from qgs_b200.integrators.integrate import rk4_tableau, n_records_of
    from qgs_b200.toolbox.lyapunov import LyapunovsEstimator, _subtimes
    t0, tw, t, dt, mdt = LYAP_ARGS
    ic = initial_conditions(1000 + rank, LYAP_MEMBERS_PER_GPU)
    N = ic.shape[0]
    b, c, a = rk4_tableau()
    pre = np.concatenate((np.arange(t0, tw, dt), np.full((1,), tw)))
    rec = np.concatenate((np.arange(tw, t, dt), np.full((1,), t)))
    ptr_a, sub_a = _subtimes(pre, mdt)
    ptr_b, sub_b = _subtimes(rec, mdt)
    sub_ptr = np.ascontiguousarray(np.concatenate((ptr_a, ptr_b[1:] + ptr_a[-1])), dtype=np.int64)
    sub_dt = np.ascontiguousarray(np.concatenate((sub_a, sub_b)))
    dt_macro = np.ascontiguousarray(np.concatenate((np.diff(pre), np.diff(rec))))
    n_pre, n_rec = len(pre) - 1, len(rec) - 1
    assert n_pre + n_rec == LYAP_STEPS
    R = n_records_of(rec, LYAP_WRITE_STEPS)
    rec_traj, rec_exp = np.empty((N, NDIM, R)), np.empty((N, LYAP_NVEC, R))
    ms = ctypes.c_double()
    _lib.set_seed(21217, rank * N)

    def device_call():
        # q0 = NULL: start bases drawn and factorised on the device; rec_vec = NULL: exponents only
        _lib.check(lib.qgsb_lyap_benettin(
            f.tensor.handle, N, _lib.dptr(ic), 0, LYAP_NVEC, None, None, n_pre, n_rec, _lib.dptr(dt_macro),
            sub_ptr.ctypes.data_as(_lib.c_long_p), _lib.dptr(sub_dt), 4, _lib.dptr(a), _lib.dptr(b), _lib.dptr(c),
            LYAP_WRITE_STEPS, 0, 1., R, _lib.dptr(rec_traj), _lib.dptr(rec_exp), None, None, None, ctypes.byref(ms)))
        return ms.value

    device_call()                                   # warm-up (module load, pool growth)
    barrier()
    launches0 = _lib.launch_count()
    dev_s = sum(device_call() for _ in range(repeats)) * 1e-3
    launches = _lib.launch_count() - launches0
    finite = bool(np.all(np.isfinite(rec_exp)))
    # end to end through the reference-facing class: host ic in, (time, traj, exponents) out
    est = LyapunovsEstimator()
    est.set_func(f, Df)

    def e2e_call():
        est.compute_lyapunovs(t0, tw, t, dt, mdt, ic, write_steps=LYAP_WRITE_STEPS, vectors=False,
                              member_offset=rank * N)
        return est.get_lyapunovs()

    e2e_call()
    barrier()
    t_wall = time.perf_counter()
    for _ in range(repeats):
        res = e2e_call()
    barrier()
    e2e_s = time.perf_counter() - t_wall
    spectrum = np.asarray(res[2]).reshape(N, LYAP_NVEC, -1).mean(axis=(0, 2))
    return dev_s, e2e_s, launches, finite, spectrum, (N * NDIM * 8, N * (NDIM + LYAP_NVEC) * R * 8)


# BASELINE.json configs[0], [2], [3]: tensor fixture, members, steps per launch, flops per member-step
SUITE = (("rp", "qgs_rp.py 2-layer channel + orography, 20 variables", 1 << 20, 500, 2620),
         ("dynT", "MAOOAM with dynamic temperatures, 38 variables, rank-5 tensor", 1 << 20, 200, 5748),
         ("T4", "MAOOAM with T^4 radiation, 38 variables, 5340 rank-5 entries", 148 * 2 * 128 * 4, 50, 103468),
         ("atm6x6", "6x6 large-basis atmosphere, 228 variables (test_aotensor_6x6)", 148 * 96 * 4, 20, 329488))


def suite_leg(lib, _lib, peak):
    """Device-timed RK4 throughput of the secondary configurations on this GPU (best of three launches)."""
    import ctypes
    from qgs_b200.functions.tendencies import tendencies_from_tensor
    from qgs_b200.integrators.integrate import rk4_tableau
    b, c, a = rk4_tableau()
    out = {}
    for name, what, members, steps, flops in SUITE:
        z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
        f, _ = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
        ens = ctypes.c_void_p()
        _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, members, ctypes.byref(ens)))
        ic = np.random.default_rng(1).random((members, f.ndim)) * (0.1 if name == "rp" else 0.01)
        _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic)))
        dt = np.full(steps, DT)
        ms = ctypes.c_double()
        best = 1e30
        for _ in range(4):
            _lib.check(lib.qgsb_ensemble_integrate(ens, steps, _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b),
                                                   _lib.dptr(c), ctypes.byref(ms)))
            best = min(best, ms.value)
        s1, s2 = np.empty(f.ndim), np.empty(f.ndim)
        _lib.check(lib.qgsb_ensemble_moments(ens, _lib.dptr(s1), _lib.dptr(s2)))
        lib.qgsb_ensemble_destroy(ens)
        rate = members * steps / (best * 1e-3)
        out[name] = {"workload": "%s; %d members x %d RK4 steps, dt=0.1, write_steps=0" % (what, members, steps),
                     "value": rate, "unit": UNIT, "ms_per_launch": best, "kernel_kind": f.tensor.kernel_kind,
                     "roofline": {"bound": "fp64", "achieved": rate * flops / 1e12, "peak": peak, "unit": "TFLOP/s",
                                  "frac": rate * flops / 1e12 / peak, "flops_per_member_step": flops},
                     "finite": bool(np.all(np.isfinite(s1)))}
        del f
    return out


def run_ours(args, rank, world, local_rank):
    import ctypes

    # host-side baselines first, on rank 0 alone and before the process group exists: the other ranks wait in the
    # rendezvous of init_process_group (a sleeping TCP wait), not in a spinning NCCL barrier
    base = port = lyap_cpu = None
    if rank == 0:
        base, port, lyap_cpu = cpu_legs(10.0)

    import torch
    from qgs_b200 import _lib
    from qgs_b200.functions.tendencies import tendencies_from_tensor
    from qgs_b200.integrators.integrate import rk4_tableau, directed_dt

    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the one
        # JSON line by pointing fd 1 at stderr until the group exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                    timeout=datetime.timedelta(minutes=30))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.init(local_rank)
    lib = _lib.load()

    z = np.load(TENSOR)
    f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
    if f.tensor.kernel_kind != 2:
        raise RuntimeError("the MAOOAM-36 specialised kernels are not linked into libqgsb.so")
    b, c, a = rk4_tableau()
    time_v = time_vector(STEPS_PER_LAUNCH)
    dt = directed_dt(time_v, 1)
    n_steps = len(dt)
    members = MEMBERS_PER_GPU

    # pinned host buffers for the end-to-end leg
    ic_host = torch.empty((members, NDIM), dtype=torch.float64).pin_memory()
    out_host = torch.empty((members, NDIM, 1), dtype=torch.float64).pin_memory()
    ic_np, out_np = ic_host.numpy(), out_host.numpy()
    ic_np[:] = initial_conditions(rank, members)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.check(lib.qgsb_synchronize())

    # ---- resident leg: `value` ----
    ens = ctypes.c_void_p()
    _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, members, ctypes.byref(ens)))
    _lib.check(lib.qgsb_ensemble_upload(ens, _lib.dptr(ic_np)))
    ms = ctypes.c_double()

    def launch(handle=None):
        _lib.check(lib.qgsb_ensemble_integrate(handle or ens, n_steps, _lib.dptr(dt), 4, _lib.dptr(a), _lib.dptr(b),
                                               _lib.dptr(c), ctypes.byref(ms)))
        return ms.value

    for _ in range(args.warmup):
        launch()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    t_wall = time.perf_counter()
    dev_ms = [launch() for _ in range(args.steps)]
    barrier()
    wall_s = time.perf_counter() - t_wall
    n_launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(dev_ms))

    # ---- end-to-end leg through the reference-facing C ABI with host buffers ----
    def e2e_call():
        _lib.check(lib.qgsb_rk_integrate(f.tensor.handle, members, _lib.dptr(ic_np), n_steps, _lib.dptr(dt), 4,
                                         _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), 0, 1, 1, _lib.dptr(out_np), None))

    for _ in range(min(args.warmup, 3)):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- ensemble statistics: the one place the path has an exchange step (NCCL all-reduce of 2n sums) ----
    s1 = np.empty(NDIM)
    s2 = np.empty(NDIM)
    _lib.check(lib.qgsb_ensemble_moments(ens, _lib.dptr(s1), _lib.dptr(s2)))
    moments = torch.from_numpy(np.concatenate((s1, s2))).cuda()
    lib.qgsb_ensemble_destroy(ens)

    # ---- strong scaling: the SAME 2^20-member ensemble split over the ranks (balanced contiguous blocks) ----
    lo, hi = rank * members // world, (rank + 1) * members // world
    strong_ms = total_ms
    if world > 1:
        ens_s = ctypes.c_void_p()
        _lib.check(lib.qgsb_ensemble_create(f.tensor.handle, hi - lo, ctypes.byref(ens_s)))
        _lib.check(lib.qgsb_ensemble_upload(ens_s, _lib.dptr(ic_np[lo:hi])))
        for _ in range(args.warmup):
            launch(ens_s)
        barrier()
        strong_ms = float(sum(launch(ens_s) for _ in range(args.steps)))
        barrier()
        lib.qgsb_ensemble_destroy(ens_s)

    # ---- Lyapunov configuration ----
    lyap_repeats = max(1, min(args.steps, 3))
    ly_dev_s, ly_e2e_s, ly_launches, ly_finite, spectrum, ly_bytes = lyapunov_leg(lib, _lib, f, Df, rank, world,
                                                                                  barrier, lyap_repeats)

    if dist is not None:
        dist.all_reduce(moments)
        worst = torch.tensor([total_ms, e2e_s, wall_s, strong_ms, ly_dev_s, ly_e2e_s], dtype=torch.float64).cuda()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, wall_s, strong_ms, ly_dev_s, ly_e2e_s = [float(v) for v in worst.cpu()]
    mean = (moments[:NDIM] / (members * world)).cpu().numpy()
    finite = bool(np.all(np.isfinite(mean)))

    if rank == 0:
        member_steps = float(members) * n_steps * args.steps * world
        value = member_steps / (total_ms * 1e-3)
        peak = _lib.fp64_peak()
        suite = suite_leg(lib, _lib, peak)
        ach = FLOPS_PER_MEMBER_STEP * float(members) * n_steps / (total_ms / args.steps * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(REPO, "profiles", "r01_rk_chain_ncu_200steps.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        ly_steps = float(LYAP_MEMBERS_PER_GPU) * LYAP_STEPS * lyap_repeats
        ly_ach = LYAP_FLOPS_PER_MEMBER_STEP * ly_steps / ly_dev_s / 1e12       # per GPU: the slowest rank's time
        peak_note = ("DFMA micro-benchmark measured on this device in this run (qgsb_fp64_peak); "
                     "MEASURED_PEAKS.json has no FP64 entry")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(world),
                "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                             "traffic": traffic, "peak_source": peak_note,
                             "flops_per_member_step": FLOPS_PER_MEMBER_STEP, "kernel": "rk_chain_kernel (maooam36)"},
                "cpu_baseline": base, "cpu_port": port,
                "e2e": {"value": member_steps / e2e_s, "unit": UNIT,
                        "h2d_bytes_per_step": world * (members * NDIM * 8 + n_steps * 8),
                        "d2h_bytes_per_step": world * members * NDIM * 8,
                        "call": "qgsb_rk_integrate (C ABI behind RungeKuttaIntegrator.integrate), pinned host buffers"},
                "strong": {"value": float(members) * n_steps * args.steps / (strong_ms * 1e-3), "unit": UNIT,
                           "members_total": members, "ms_per_step": strong_ms / args.steps,
                           "note": "the same 2^20-member ensemble split over the %d GPU(s); device-timed, max over "
                                   "ranks" % world},
                "lyapunov": {
                    "metric": LYAP_METRIC, "value": ly_steps * world / ly_dev_s, "unit": UNIT,
                    "ms_per_step": ly_dev_s / lyap_repeats * 1e3, "steps": lyap_repeats, "scaling": "weak",
                    "config": {"workload": "LyapunovsEstimator.compute_lyapunovs(0, 100, 200, 0.1, 0.1, write_steps=10), "
                                           "MAOOAM-36, %d members/GPU, 36 vectors, mdt = dt: %d Benettin steps per "
                                           "member and call; start bases drawn on the device; exponents only"
                                           % (LYAP_MEMBERS_PER_GPU, LYAP_STEPS),
                               "members_per_gpu": LYAP_MEMBERS_PER_GPU, "n_vec": LYAP_NVEC},
                    "roofline": {"bound": "fp64", "achieved": ly_ach, "peak": peak, "unit": "TFLOP/s",
                                 "frac": ly_ach / peak, "traffic": None, "peak_source": peak_note,
                                 "flops_per_member_step": LYAP_FLOPS_PER_MEMBER_STEP,
                                 "kernel": "pack::lyap_kernel<36, bilinear product; Cholesky QR on mma.sync.m8n8k4.f64 between records, Householder at observed steps>"},
                    "e2e": {"value": ly_steps * world / ly_e2e_s, "unit": UNIT,
                            "h2d_bytes_per_step": world * ly_bytes[0], "d2h_bytes_per_step": world * ly_bytes[1],
                            "call": "LyapunovsEstimator.compute_lyapunovs(..., vectors=False) + get_lyapunovs(): host "
                                    "numpy arrays in and out"},
                    "cpu_baseline": lyap_cpu, "gpu_launches": int(ly_launches), "exponents_finite": ly_finite,
                    "leading_exponents": [float(v) for v in spectrum[:4]]},
                "suite": suite,
                "gpu_launches": int(n_launches), "clocks": clocks,
                "wall_s_timed_region": wall_s, "ensemble_mean_finite": finite}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if world == 1 and args.gpus > 1:
        # not under torchrun: launch one rank per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
