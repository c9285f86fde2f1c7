/*
 * qgsb.h -- C ABI of libqgsb.so, the B200 (sm_100a) ensemble integrator for the qgs hot path.
 *
 * Every entry point replaces one numba-jitted function (or worker pool) of the reference; the
 * reference file:line it stands in for is cited next to each declaration (paths relative to the
 * reference repository root).  The reference-side binding a maintainer would add is the ctypes stub
 * shown in INTEGRATION.md; qgs_b200/_lib.py is that stub.
 *
 * Conventions
 *   - plain C types only; all array arguments are HOST pointers owned by the caller for the duration
 *     of the call unless the name starts with d_ (device pointer) -- the library owns every device
 *     allocation it makes (tensor handles, ensembles, scratch).
 *   - arithmetic is IEEE float64; indices are int32 (the Python side converts the reference's
 *     F-ordered int64 `coords.T` view, tendencies.py:92, to C-ordered int32).
 *   - return value 0 = success; non-zero = failure, message in qgsb_last_error() (thread local).
 *   - calls are synchronous (they return when results are in the caller's buffers), like the
 *     reference's integrate() which blocks on queue.join() (integrator.py:391).  The library may be
 *     called from several host threads: entry points serialise on one process-wide lock.
 *   - devices: a process drives ONE device (one process per GPU under torchrun: LOCAL_RANK) or SEVERAL (any other
 *     process on a multi-GPU box, see qgsb_init).  With several, the host-buffer entry points qgsb_rk_integrate,
 *     qgsb_rk_tgls_integrate, qgsb_lyap_benettin and qgsb_clv_ginelli split the members into contiguous blocks
 *     [g N / G, (g + 1) N / G), one per device, and run them concurrently on the devices' own streams -- the
 *     reference's fan-out of one integrate() over num_threads worker processes (integrator.py:121-142, 386-395).
 *     Members are independent, nothing is exchanged, results are bitwise those of one device.  Tensor handles
 *     live on the primary device (the first of the list; the library replicates them where needed); a large resident
 *     ensemble is spread over the devices like the host-buffer calls are.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with an error.
 */
#ifndef QGSB_H
#define QGSB_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define QGSB_API __attribute__((visibility("default")))
#else
#define QGSB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qgsb_tensor qgsb_tensor;     /* device-resident tendencies + Jacobian tensor */
typedef struct qgsb_ensemble qgsb_ensemble; /* device-resident ensemble state (structure of arrays) */

/* ---- runtime ---------------------------------------------------------------------------------- */

/* Choose the device(s) and create the library streams.  Replaces RungeKuttaIntegrator.start() spawning the worker
 * pool (qgs/integrators/integrator.py:121-142).  device >= 0: that device only.  device < 0: keep the current
 * configuration when there is one; else QGSB_DEVICES ("all", a count, or a comma-separated list of ordinals) decides;
 * without it a torchrun rank (LOCAL_RANK set) takes its own device and any other process every visible device.
 * Idempotent. */
QGSB_API int qgsb_init(int device);
/* Drive exactly these devices (CUDA ordinals; the first is the primary one).  Handles created before stay valid as
 * long as their device is still visible; resident ensembles must be re-created when the primary device changes. */
QGSB_API int qgsb_set_devices(int n_devices, const int *devices);
/* Number of devices the library drives (0 before initialisation). */
QGSB_API int qgsb_device_count(void);
/* Release scratch memory and the stream; replaces terminate() (integrator.py:113-119).  Idempotent. */
QGSB_API void qgsb_shutdown(void);
/* Launch on a caller-provided cudaStream_t (e.g. torch's current stream); NULL restores the own stream. */
QGSB_API int qgsb_set_stream(void *cuda_stream);
QGSB_API const char *qgsb_last_error(void);
QGSB_API const char *qgsb_version(void);
QGSB_API int qgsb_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
/* Number of kernels this library has launched since load (bench.py's "gpu_launches"). */
QGSB_API long qgsb_launch_count(void);
/* Load a shared object holding tensor-specialised kernels built by qgs_b200/codegen.py; handles
 * created earlier pick it up through qgsb_tensor_use_specialised(t, 1). */
QGSB_API int qgsb_load_plugin(const char *path);

/* ---- tensor hand-off (qgs/functions/tendencies.py:92-96) ---------------------------------------
 * coo (nnz, rank) row-major, val (nnz); jcoo/jval the Jacobian tensor (may be NULL/0).  rank is 3
 * (sparse_mul3/sparse_mul2) or 5 (sparse_mul5/sparse_mul4).  Index 0 is the constant x_0 = 1. */
QGSB_API int qgsb_tensor_create(int ndim, int rank, long nnz, const int32_t *coo, const double *val,
                       long jnnz, const int32_t *jcoo, const double *jval, qgsb_tensor **out);
QGSB_API void qgsb_tensor_destroy(qgsb_tensor *t);
/* kernel_kind: 0 generic thread-per-member, 1 generic warp-per-member, 2 tensor-specialised.
 * tensor_hash: FNV-1a over (ndim, rank, nnz, row-sorted coo, values), the key of a specialised module. */
QGSB_API int qgsb_tensor_info(const qgsb_tensor *t, int *ndim, int *rank, long *nnz, long *jnnz, int *kernel_kind,
                     uint64_t *tensor_hash);
/* Host-only: the key qgsb_tensor_info reports, computed without touching the device (coo is sorted by
 * row internally).  Mirrors qgs_b200.codegen.tensor_hash. */
QGSB_API int qgsb_tensor_hash(int ndim, int rank, long nnz, const int32_t *coo, const double *val, uint64_t *hash);
/* Forbid / allow the tensor-specialised kernels for this handle (testing and benchmarking). */
/* 1 when the handle runs generated tangent-linear / Benettin kernels (module loaded and its Jacobian tensor matches) */
QGSB_API int qgsb_tensor_has_tangent(const qgsb_tensor *t);
QGSB_API int qgsb_tensor_use_specialised(qgsb_tensor *t, int enable);

/* ---- raw contractions, qgs/functions/sparse_mul.py ---------------------------------------------- */
/* sparse_mul3 (sparse_mul.py:48-81): res_i = sum T_ijk a_j b_k, res_0 = 1.        vectors length n1 */
QGSB_API int qgsb_sparse_mul3(long nnz, const int32_t *coo, const double *val, int n1,
                     const double *vec_a, const double *vec_b, double *res);
/* sparse_mul5 (sparse_mul.py:121-158) */
QGSB_API int qgsb_sparse_mul5(long nnz, const int32_t *coo, const double *val, int n1, const double *vec_a,
                     const double *vec_b, const double *vec_c, const double *vec_d, double *res);
/* sparse_mul2 (sparse_mul.py:13-45): res_ij = sum_k T_ijk a_k, res is (n1, n1) */
QGSB_API int qgsb_sparse_mul2(long nnz, const int32_t *coo, const double *val, int n1, const double *vec, double *res);
/* sparse_mul4 (sparse_mul.py:84-118) */
QGSB_API int qgsb_sparse_mul4(long nnz, const int32_t *coo, const double *val, int n1, const double *vec_a,
                     const double *vec_b, const double *vec_c, double *res);

/* ---- f / Df closures, qgs/functions/tendencies.py:98-121, batched over members ------------------ */
QGSB_API int qgsb_tendencies(const qgsb_tensor *t, long n_traj, const double *x /* (N, n) */, double *out /* (N, n) */);
QGSB_API int qgsb_jacobian(const qgsb_tensor *t, long n_traj, const double *x /* (N, n) */, double *out /* (N, n, n) */);

/* ---- _integrate_runge_kutta_jit, qgs/integrators/integrate.py:182-223 ---------------------------
 * dt (n_steps): the signed step lengths in execution order, i.e. np.diff(directed_time).  The model
 * is autonomous (tendencies.py:99 ignores t) so the times themselves are not needed.
 * a (s, s) row-major, b (s), c (s) the Butcher tableau (explicit part j < i is used; c is accepted
 * for signature parity).  write_steps / n_records follow integrate.py:190-196, 210-212, 221;
 * time_direction -1 reverses the record axis like integrate.py:223.
 * traj (n_traj, n_dim, n_records) C-ordered.  device_ms (may be NULL): kernel time by CUDA events. */
QGSB_API int qgsb_rk_integrate(const qgsb_tensor *t, long n_traj, const double *ic /* (N, n) */, long n_steps,
                      const double *dt, int s, const double *a, const double *b, const double *c,
                      long write_steps, int time_direction, long n_records, double *traj,
                      double *device_ms);

/* ---- _integrate_runge_kutta_tgls_jit, integrate.py:555-614 (+ :226-231) ---------------------------
 * tg_ic (N, n, m); adjoint != 0 integrates with J^T; inverse_sign is the reference's `inverse`
 * (+1. or -1., integrate.py:519-521); the inhomogeneous term is the reference default _zeros_func.
 * fmat (N, n, m, n_records). */
QGSB_API int qgsb_rk_tgls_integrate(const qgsb_tensor *t, long n_traj, const double *ic, int m, const double *tg_ic,
                           long n_steps, const double *dt, int s, const double *a, const double *b,
                           const double *c, long write_steps, int time_direction, int adjoint,
                           double inverse_sign, long n_records, double *traj, double *fmat,
                           double *device_ms);

/* ---- Benettin steps, qgs/toolbox/lyapunov.py:471-632 ----------------------------------------------
 * One call runs n_pre + n_rec macro steps for every member.  Macro step i advances the nonlinear
 * state with dt_macro[i] (the stored write_steps=1 trajectory of lyapunov.py:558 / :474) and the
 * n x n_vec tangent matrix with the micro steps sub_dt[sub_ptr[i] .. sub_ptr[i+1]) starting from the
 * same point (lyapunov.py:598-601), then re-orthonormalises by Householder QR (np.linalg.qr, :602).
 * forward == 0 (BLV, :564-632): members start at ic; records taken during the last n_rec steps.
 * forward == 1 (FLV, :480-552): the trajectory is first integrated forward over dt_macro reversed and
 *   stored in HBM, then walked backwards; dt_macro / sub_dt are given in execution (backward) order, negative.
 * forward == 2: like 0 but the nonlinear state follows the micro steps (Ginelli forward pass, :1212-1218).
 * q0 (N, n, n_vec), r0 (N, n_vec, n_vec) or NULL: the start basis.  The reference draws
 * qr(random((n_dim, n_vec))), lyapunov.py:592-593; q0 == NULL does the same ON THE DEVICE (counter-based uniform
 * draw keyed by qgsb_set_seed and the member's index, factorised by the kernel's own Householder QR), so no
 * (N, n, n_vec) array crosses the bus and no host LAPACK runs.
 * rec_* follow _compute_*_lyap_traj_jit's return values: traj (N, n, R), exp (N, n_vec, R),
 * vec (N, n, n_vec, R); rec_vec may be NULL when only the exponents are wanted (no vector records are kept
 * or copied).  r_all (N, n_steps_total, n_vec, n_vec) or NULL stores every R factor and
 * q_all (N, n_rec + 1, n, n_vec) or NULL the basis at every recorded-phase point (Ginelli, :1220-1250). */
QGSB_API int qgsb_lyap_benettin(const qgsb_tensor *t, long n_traj, const double *ic, int forward, int n_vec,
                       const double *q0, const double *r0, long n_pre, long n_rec,
                       const double *dt_macro, const long *sub_ptr, const double *sub_dt,
                       int s, const double *a, const double *b, const double *c, long write_steps,
                       int adjoint, double inverse_sign, long n_records,
                       double *rec_traj, double *rec_exp, double *rec_vec,
                       double *r_all, double *q_all, double *device_ms);

/* Seed of the device-side start bases (q0 == NULL above); member_offset is added to the member index, so that a
 * process holding members [lo, hi) of a larger ensemble draws what a single process would have drawn for them. */
QGSB_API int qgsb_set_seed(uint64_t seed, long member_offset);

/* Covariant Lyapunov vectors, method of Ginelli et al. -- replaces _compute_clv_gin_jit
 * (qgs/toolbox/lyapunov.py:1174-1288) with solve_triangular_matrix / normalize_matrix_columns
 * (qgs/functions/util.py:56-98).  Forward Benettin pass over n_pre (t0..ta), n_time (ta..tb) and n_after (tb..tc)
 * steps (dt_macro / sub_ptr / sub_dt as for qgsb_lyap_benettin, n_pre + n_time + n_after steps) that keeps every Q
 * and R of [ta, tc] in HBM, then the backward recursion A <- normalise(R^-1 A + noise_pert * diag(noise)) from tc
 * to ta starting from am0.  Records follow the time vector ta..tb: n_records of them, the last written first.
 * am0: host (N, m, m) upper triangular with unit columns; noise: host (N, n_time + n_after, m) or NULL when
 * noise_pert == 0; dte: host (n_time + 1) step lengths used for the local exponents.  Members are processed in
 * batches sized to the device memory. */
QGSB_API int qgsb_clv_ginelli(const qgsb_tensor *t, long n_traj, const double *ic, int n_vec, const double *q0,
                              const double *r0, long n_pre, long n_time, long n_after, const double *dt_macro,
                              const long *sub_ptr, const double *sub_dt, int s, const double *a, const double *b,
                              const double *c, long write_steps, double noise_pert, const double *am0,
                              const double *noise, const double *dte, long n_records,
                              double *rec_traj /* (N, n, R) */, double *rec_exp /* (N, m, R) */,
                              double *rec_vec /* (N, n, m, R) */, double *device_ms);

/* Covariant Lyapunov vectors by subspace intersection -- replaces the per-record SVD loop of _compute_clv_sub_jit
 * (qgs/toolbox/lyapunov.py:1315-1320): CLV j = the direction common to span(BLV_0..j) and span(FLV_0..n-1-j), found
 * as the null vector of a corner of FLV^T BLV instead of as a first singular vector (same direction; its sign, which
 * is arbitrary in the reference too, is fixed by making the largest coefficient positive).  bvec, fvec, vec: host
 * arrays (n_traj, n_dim, n_dim, n_records) in the layout of get_blvs() / get_flvs() / get_clvs(). */
QGSB_API int qgsb_clv_subspace_intersect(long n_traj, int n_dim, long n_records, const double *bvec,
                                         const double *fvec, double *vec);

/* ---- device-resident ensemble (state stays in HBM between calls) ---------
 * State layout in HBM: tiled structure of arrays -- members in tiles of 128, variable i of member m at
 * (m / 128) * (n * 128) + i * 128 + m % 128; ld = n_traj rounded up to 128. */
QGSB_API int qgsb_ensemble_create(const qgsb_tensor *t, long n_traj, qgsb_ensemble **out);
QGSB_API void qgsb_ensemble_destroy(qgsb_ensemble *e);
QGSB_API int qgsb_ensemble_upload(qgsb_ensemble *e, const double *ic /* host (N, n) */);
QGSB_API int qgsb_ensemble_download(qgsb_ensemble *e, double *out /* host (N, n) */);
/* Advance all members n_steps (write_steps = 0 semantics: only the end state is kept).  One fused
 * launch; returns without host synchronisation when device_ms is NULL. */
QGSB_API int qgsb_ensemble_integrate(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                            const double *b, const double *c, double *device_ms);
/* Same, recording like qgsb_rk_integrate into a device buffer d_rec of n_records tiled-SoA states
 * (record r starts at r * n * ld). */
QGSB_API int qgsb_ensemble_integrate_record(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                   const double *b, const double *c, long write_steps, long n_records,
                                   double *d_rec, double *device_ms);
/* Sum and sum of squares over members for every variable (ensemble statistics; NCCL all-reduce of
 * these 2n doubles is done by the caller across ranks). */
/* Same integration, but the records stream to the caller's host array traj (N, n, n_records) -- the layout of
 * qgsb_rk_integrate / get_trajectories() -- in chunks while the next chunk integrates; device memory is bounded
 * independently of n_records and the ensemble stays resident (qgs_maooam.py:115-136 with write_steps > 0). */
QGSB_API int qgsb_ensemble_integrate_trajectories(qgsb_ensemble *e, long n_steps, const double *dt, int s,
                                                  const double *a, const double *b, const double *c,
                                                  long write_steps, int time_direction, long n_records,
                                                  double *traj, double *device_ms);
QGSB_API int qgsb_ensemble_moments(qgsb_ensemble *e, double *sum /* (n) */, double *sumsq /* (n) */);
/* Integrate the resident ensemble and return, for every record of integrate.py:190-221 (n_records of them), the
 * per-variable sum and sum of squares over the members -- the ensemble statistics of
 * qgs/integrators/statistics.py:33-66 (TrajectoriesStatistics.compute_stats with the identity and the square as
 * functions) without materialising the (n_traj, n_dim, n_records) trajectories on the host.  The ensemble ends at
 * the final state, like qgsb_ensemble_integrate. */
QGSB_API int qgsb_ensemble_integrate_moments(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                             const double *b, const double *c, long write_steps, long n_records,
                                             double *sum /* host (R, n) */, double *sumsq /* host (R, n) */,
                                             double *device_ms);
/* Device pointer / leading dimension of the state -- for an ensemble that lives on ONE device; NULL / 0 for an ensemble
 * spread over several (qgsb_ensemble_integrate_record, which takes a device record buffer, refuses those too). */
QGSB_API void *qgsb_ensemble_device_ptr(qgsb_ensemble *e);
QGSB_API long qgsb_ensemble_ld(const qgsb_ensemble *e);
QGSB_API int qgsb_synchronize(void);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* Dependent-chain-free DFMA micro-benchmark: measured FP64 FMA peak of this device in TFLOP/s
 * (the roofline denominator MEASURED_PEAKS.json lacks). */
QGSB_API int qgsb_fp64_peak(double *tflops, double *ms);
/* Same for the FP64 tensor-core path (mma.sync m8n8k4 f64). */
QGSB_API int qgsb_dmma_peak(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* QGSB_H */
