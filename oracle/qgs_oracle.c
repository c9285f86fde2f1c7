/*
 * qgs_oracle.c -- CPU restatement of the qgs ensemble-integrator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under qgs_b200/ may import, link or call this file; it is
 * used by tests/, by __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
 * --impl reference leg as the CPU arm ("port").  The product path is the CUDA library.
 *
 * Parity status: PINNED.  The reference holds no golden value of f(x), of a trajectory or of a
 * Lyapunov exponent, so this file is pinned against outputs of the unmodified
 * reference numba code generated in the build container by tests/golden/make_golden.py and
 * committed under tests/golden/ (see tests/test_oracle.py), and its input tensors are pinned
 * against the reference's own model_test/test_aotensor*.ref golden files.
 *
 * Every function cites the reference lines (relative to /root/reference) whose arithmetic it
 * follows.  Operation order inside a term matches the reference ((a*b)*value, then +=) and the file
 * is compiled with -ffp-contract=off so that no FMA is formed, like numba's default code.
 *
 * Layouts are the reference's: coo is (nnz, rank) row-major int32, states are C-ordered doubles,
 * trajectories are (n_traj, n_dim, n_records), fundamental matrices (n_traj, n_dim, n_tg, n_records).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for over ensemble members (the reference fans members over
 *      multiprocessing workers, integrator.py:121-142, 388-395) ---- */
static int g_threads = 0;
typedef void (*range_fn)(void *ctx, long begin, long end);
typedef struct { range_fn fn; void *ctx; long begin, end; } range_job;
static void *range_tramp(void *p) { range_job *j = (range_job *)p; j->fn(j->ctx, j->begin, j->end); return NULL; }
static void parallel_for(long N, range_fn fn, void *ctx)
{
    int nt = g_threads > 0 ? g_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt > N) nt = (int)N;
    if (nt <= 1) { fn(ctx, 0, N); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    range_job *jobs = (range_job *)malloc(sizeof(range_job) * nt);
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].begin = N * t / nt; jobs[t].end = N * (t + 1) / nt;
        pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct {
    int ndim;            /* number of model variables n; vectors below have n+1 entries (x_0 = 1) */
    int rank;            /* 3 or 5 */
    long nnz;
    const int32_t *coo;  /* (nnz, rank) */
    const double *val;
    long jnnz;
    const int32_t *jcoo; /* (jnnz, rank) */
    const double *jval;
} qgso_tensor;

/* ---------------------------------------------------------------------------------------------
 * qgs/functions/sparse_mul.py
 * ------------------------------------------------------------------------------------------ */

/* sparse_mul.py:13-45   A_ij = sum_k T_ijk a_k,   res is (n1, n1) */
void qgso_sparse_mul2(long nnz, const int32_t *coo, const double *val, int n1, const double *vec, double *res)
{
    memset(res, 0, sizeof(double) * (size_t)n1 * n1);
    for (long e = 0; e < nnz; ++e) {
        const int32_t *c = coo + 3 * e;
        res[(size_t)c[0] * n1 + c[1]] += vec[c[2]] * val[e];
    }
}

/* sparse_mul.py:48-81   v_i = sum_jk T_ijk a_j b_k,  v_0 = 1 */
void qgso_sparse_mul3(long nnz, const int32_t *coo, const double *val, int n1,
                      const double *va, const double *vb, double *res)
{
    memset(res, 0, sizeof(double) * (size_t)n1);
    for (long e = 0; e < nnz; ++e) {
        const int32_t *c = coo + 3 * e;
        res[c[0]] += va[c[1]] * vb[c[2]] * val[e];
    }
    res[0] = 1.;
}

/* sparse_mul.py:84-118  A_ij = sum_klm T_ijklm a_k b_l c_m */
void qgso_sparse_mul4(long nnz, const int32_t *coo, const double *val, int n1,
                      const double *va, const double *vb, const double *vc, double *res)
{
    memset(res, 0, sizeof(double) * (size_t)n1 * n1);
    for (long e = 0; e < nnz; ++e) {
        const int32_t *c = coo + 5 * e;
        res[(size_t)c[0] * n1 + c[1]] += va[c[2]] * vb[c[3]] * vc[c[4]] * val[e];
    }
}

/* sparse_mul.py:121-158 v_i = sum_jklm T_ijklm a_j b_k c_l d_m,  v_0 = 1 */
void qgso_sparse_mul5(long nnz, const int32_t *coo, const double *val, int n1,
                      const double *va, const double *vb, const double *vc, const double *vd, double *res)
{
    memset(res, 0, sizeof(double) * (size_t)n1);
    for (long e = 0; e < nnz; ++e) {
        const int32_t *c = coo + 5 * e;
        res[c[0]] += va[c[1]] * vb[c[2]] * vc[c[3]] * vd[c[4]] * val[e];
    }
    res[0] = 1.;
}

/* ---------------------------------------------------------------------------------------------
 * qgs/functions/tendencies.py:98-121 -- the f / Df closures.  `work` holds 2(n+1) doubles for f,
 * (n+1) + (n+1)^2 for Df.
 * ------------------------------------------------------------------------------------------ */
static void f_eval(const qgso_tensor *T, const double *x, double *out, double *work)
{
    const int n = T->ndim, n1 = n + 1;
    double *xx = work, *xr = work + n1;
    xx[0] = 1.;
    memcpy(xx + 1, x, sizeof(double) * n);
    if (T->rank == 5)
        qgso_sparse_mul5(T->nnz, T->coo, T->val, n1, xx, xx, xx, xx, xr);
    else
        qgso_sparse_mul3(T->nnz, T->coo, T->val, n1, xx, xx, xr);
    memcpy(out, xr + 1, sizeof(double) * n);
}

static void jac_eval(const qgso_tensor *T, const double *x, double *out /* n x n */, double *work)
{
    const int n = T->ndim, n1 = n + 1;
    double *xx = work, *full = work + n1;
    xx[0] = 1.;
    memcpy(xx + 1, x, sizeof(double) * n);
    if (T->rank == 5)
        qgso_sparse_mul4(T->jnnz, T->jcoo, T->jval, n1, xx, xx, xx, full);
    else
        qgso_sparse_mul2(T->jnnz, T->jcoo, T->jval, n1, xx, full);
    for (int i = 0; i < n; ++i)
        memcpy(out + (size_t)i * n, full + (size_t)(i + 1) * n1 + 1, sizeof(double) * n);
}

/* batched public forms: x (N, n) -> out (N, n) / (N, n, n) */
typedef struct { const qgso_tensor *T; const double *x; double *out; } tend_ctx;

static void tend_range(void *p, long b, long e)
{
    tend_ctx *c = (tend_ctx *)p;
    const int n = c->T->ndim;
    double *work = (double *)malloc(sizeof(double) * 2 * (n + 1));
    for (long m = b; m < e; ++m) f_eval(c->T, c->x + m * n, c->out + m * n, work);
    free(work);
}

static void jac_range(void *p, long b, long e)
{
    tend_ctx *c = (tend_ctx *)p;
    const int n = c->T->ndim;
    double *work = (double *)malloc(sizeof(double) * ((n + 1) + (size_t)(n + 1) * (n + 1)));
    for (long m = b; m < e; ++m) jac_eval(c->T, c->x + m * n, c->out + (size_t)m * n * n, work);
    free(work);
}

void qgso_tendencies(const qgso_tensor *T, long N, const double *x, double *out)
{
    tend_ctx c = {T, x, out};
    parallel_for(N, tend_range, &c);
}

void qgso_jacobian(const qgso_tensor *T, long N, const double *x, double *out)
{
    tend_ctx c = {T, x, out};
    parallel_for(N, jac_range, &c);
}

/* ---------------------------------------------------------------------------------------------
 * Record bookkeeping, integrate.py:190-196:  R = len(time[::ws]) (+1 if that misses the last time)
 * ------------------------------------------------------------------------------------------ */
long qgso_n_records(long L, long ws)
{
    if (ws == 0) return 1;
    long r = (L + ws - 1) / ws;          /* len(time[::ws]) */
    if ((r - 1) * ws != L - 1) r += 1;   /* tot[-1] != time[-1]  (time is strictly monotone) */
    return r;
}

/* ---------------------------------------------------------------------------------------------
 * qgs/integrators/integrate.py:182-223  _integrate_runge_kutta_jit
 *
 * time (L,) is the forward time vector; time_direction -1 integrates it reversed (negative steps,
 * integrate.py:199-203) and the record axis is flipped at the end (:223).
 * The system is autonomous (tendencies.py:99: t unused), so only dt = diff(directed_time) matters.
 * ------------------------------------------------------------------------------------------ */
static void rk_one(const qgso_tensor *T, const double *dtime, long L, const double *ic, long ws,
                   int s, const double *a, const double *b, long R, int flip,
                   double *rec /* (n, R) strides: R,1 */, double *buf)
{
    const int n = T->ndim;
    double *y = buf, *ys = y + n, *k = ys + n, *work = k + (size_t)s * n;
    memcpy(y, ic, sizeof(double) * n);
    long iw = 0;
    for (long ti = 0; ti + 1 < L; ++ti) {
        const double dt = dtime[ti + 1] - dtime[ti];
        if (ws > 0 && ti % ws == 0) {
            long col = flip ? R - 1 - iw : iw;
            for (int d = 0; d < n; ++d) rec[(size_t)d * R + col] = y[d];
            ++iw;
        }
        memset(k, 0, sizeof(double) * (size_t)s * n);
        for (int i = 0; i < s; ++i) {
            /* y_s = y + (dt * a[i]) @ k      integrate.py:216 */
            for (int d = 0; d < n; ++d) {
                double acc = 0.;
                for (int j = 0; j < s; ++j) acc += (dt * a[i * s + j]) * k[(size_t)j * n + d];
                ys[d] = y[d] + acc;
            }
            f_eval(T, ys, k + (size_t)i * n, work);          /* :217 */
        }
        for (int d = 0; d < n; ++d) {                         /* y + (dt * b) @ k   :218 */
            double acc = 0.;
            for (int j = 0; j < s; ++j) acc += (dt * b[j]) * k[(size_t)j * n + d];
            y[d] = y[d] + acc;
        }
    }
    {
        long col = flip ? 0 : R - 1;                          /* :221 then :223 */
        for (int d = 0; d < n; ++d) rec[(size_t)d * R + col] = y[d];
    }
}

static double *directed(const double *time, long L, int dir)
{
    double *d = (double *)malloc(sizeof(double) * L);
    for (long i = 0; i < L; ++i) d[i] = dir == -1 ? time[L - 1 - i] : time[i];
    return d;
}

typedef struct {
    const qgso_tensor *T; const double *dtime; long L; const double *ic; long ws; int s;
    const double *a, *b; long R; int flip; double *traj;
} rk_ctx;

static void rk_range(void *p, long b0, long e0)
{
    rk_ctx *c = (rk_ctx *)p;
    const int n = c->T->ndim;
    double *buf = (double *)malloc(sizeof(double) * ((size_t)(2 + c->s) * n + 2 * (n + 1)));
    for (long m = b0; m < e0; ++m)
        rk_one(c->T, c->dtime, c->L, c->ic + m * n, c->ws, c->s, c->a, c->b, c->R, c->flip,
               c->traj + (size_t)m * n * c->R, buf);
    free(buf);
}

int qgso_rk_integrate(const qgso_tensor *T, long N, const double *ic, long L, const double *time,
                      int time_direction, long ws, int s, const double *a, const double *b,
                      const double *c, double *traj /* (N, n, R) */)
{
    (void)c;
    const int n = T->ndim;
    const long R = qgso_n_records(L, ws);
    double *dtime = directed(time, L, time_direction);
    memset(traj, 0, sizeof(double) * (size_t)N * n * R);
    rk_ctx cx = {T, dtime, L, ic, ws, s, a, b, R, time_direction == -1, traj};
    parallel_for(N, rk_range, &cx);
    free(dtime);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * qgs/integrators/integrate.py:555-614  _integrate_runge_kutta_tgls_jit
 * with _tangent_linear_system (:226-231) and boundary == _zeros_func (:235-237).
 *
 * tg (n, m) row-major.  km_i = inverse * (J(y_s) or J(y_s)^T) @ (fm + sum_j dt a_ij km_j).
 * When rec_y / rec_fm are NULL only the end state is returned in y_out / fm_out.
 * ------------------------------------------------------------------------------------------ */
static void tgls_one(const qgso_tensor *T, const double *dtime, long L, const double *ic, int m,
                     const double *tg_ic, long ws, int s, const double *a, const double *b,
                     int adjoint, double inverse, long R, int flip,
                     double *rec_y /* (n,R) */, double *rec_fm /* (n,m,R) */,
                     double *y_out, double *fm_out, double *buf)
{
    const int n = T->ndim;
    const size_t nm = (size_t)n * m;
    double *y = buf, *ys = y + n, *k = ys + n, *fm = k + (size_t)s * n, *kms = fm + nm,
           *km = kms + nm, *J = km + (size_t)s * nm, *work = J + (size_t)n * n;
    memcpy(y, ic, sizeof(double) * n);
    memcpy(fm, tg_ic, sizeof(double) * nm);
    long iw = 0;
    for (long ti = 0; ti + 1 < L; ++ti) {
        const double dt = dtime[ti + 1] - dtime[ti];
        if (rec_y && ws > 0 && ti % ws == 0) {
            long col = flip ? R - 1 - iw : iw;
            for (int d = 0; d < n; ++d) rec_y[(size_t)d * R + col] = y[d];
            for (size_t q = 0; q < nm; ++q) rec_fm[q * R + col] = fm[q];
            ++iw;
        }
        memset(k, 0, sizeof(double) * (size_t)s * n);
        memset(km, 0, sizeof(double) * (size_t)s * nm);
        for (int i = 0; i < s; ++i) {
            for (int d = 0; d < n; ++d) {
                double acc = 0.;
                for (int j = 0; j < s; ++j) acc += (dt * a[i * s + j]) * k[(size_t)j * n + d];
                ys[d] = y[d] + acc;
            }
            f_eval(T, ys, k + (size_t)i * n, work);
            memcpy(kms, fm, sizeof(double) * nm);                        /* :598-600 */
            for (int j = 0; j < s; ++j) {
                const double w = dt * a[i * s + j];
                for (size_t q = 0; q < nm; ++q) kms[q] += w * km[(size_t)j * nm + q];
            }
            jac_eval(T, ys, J, work);                                    /* :601, :226-231 */
            double *kmi = km + (size_t)i * nm;
            for (int r = 0; r < n; ++r)
                for (int col = 0; col < m; ++col) {
                    double acc = 0.;
                    for (int q = 0; q < n; ++q)
                        acc += (adjoint ? J[(size_t)q * n + r] : J[(size_t)r * n + q]) * kms[(size_t)q * m + col];
                    kmi[(size_t)r * m + col] = inverse * acc;            /* + boundary == 0   :602-603 */
                }
        }
        for (int d = 0; d < n; ++d) {
            double acc = 0.;
            for (int j = 0; j < s; ++j) acc += (dt * b[j]) * k[(size_t)j * n + d];
            y[d] = y[d] + acc;
        }
        for (int j = 0; j < s; ++j) {                                    /* :605-607 */
            const double w = dt * b[j];
            for (size_t q = 0; q < nm; ++q) fm[q] += w * km[(size_t)j * nm + q];
        }
    }
    if (rec_y) {
        long col = flip ? 0 : R - 1;
        for (int d = 0; d < n; ++d) rec_y[(size_t)d * R + col] = y[d];
        for (size_t q = 0; q < nm; ++q) rec_fm[q * R + col] = fm[q];
    }
    if (y_out) memcpy(y_out, y, sizeof(double) * n);
    if (fm_out) memcpy(fm_out, fm, sizeof(double) * nm);
}

static size_t tgls_buf_doubles(int n, int m, int s)
{
    return (size_t)(2 + s) * n + (size_t)(2 + s) * n * m + (size_t)n * n + (n + 1) + (size_t)(n + 1) * (n + 1);
}

typedef struct {
    const qgso_tensor *T; const double *dtime; long L; const double *ic; int m; const double *tg_ic;
    long ws; int s; const double *a, *b; int adjoint; double inverse; long R; int flip;
    double *traj, *fmat;
} tgls_ctx;

static void tgls_range(void *p, long b0, long e0)
{
    tgls_ctx *c = (tgls_ctx *)p;
    const int n = c->T->ndim, m = c->m;
    double *buf = (double *)malloc(sizeof(double) * tgls_buf_doubles(n, m, c->s));
    for (long q = b0; q < e0; ++q)
        tgls_one(c->T, c->dtime, c->L, c->ic + q * n, m, c->tg_ic + (size_t)q * n * m, c->ws, c->s,
                 c->a, c->b, c->adjoint, c->inverse, c->R, c->flip,
                 c->traj + (size_t)q * n * c->R, c->fmat + (size_t)q * n * m * c->R, NULL, NULL, buf);
    free(buf);
}

int qgso_rk_tgls_integrate(const qgso_tensor *T, long N, const double *ic, int m,
                           const double *tg_ic /* (N, n, m) */, long L, const double *time,
                           int time_direction, long ws, int s, const double *a, const double *b,
                           const double *c, int adjoint, double inverse,
                           double *traj /* (N,n,R) */, double *fmat /* (N,n,m,R) */)
{
    (void)c;
    const int n = T->ndim;
    const long R = qgso_n_records(L, ws);
    double *dtime = directed(time, L, time_direction);
    tgls_ctx cx = {T, dtime, L, ic, m, tg_ic, ws, s, a, b, adjoint, inverse, R, time_direction == -1, traj, fmat};
    (void)n;
    parallel_for(N, tgls_range, &cx);
    free(dtime);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Householder QR with LAPACK's conventions (dgeqr2 + dorg2r), which is what np.linalg.qr runs for
 * the reference (lyapunov.py:602-604): A (n, m) row-major, n >= m -> Q (n, m), R (m, m).
 * R_jj = -sign(alpha) * ||x||, so signs of Q columns match numpy's.
 * ------------------------------------------------------------------------------------------ */
void qgso_qr(int n, int m, const double *A, double *Q, double *R, double *work /* n*m + m */)
{
    double *H = work;            /* (n, m) column reflectors stored below the diagonal */
    double *tau = work + (size_t)n * m;
    memcpy(H, A, sizeof(double) * (size_t)n * m);
    for (int j = 0; j < m; ++j) {
        double alpha = H[(size_t)j * m + j], xnorm = 0.;
        for (int i = j + 1; i < n; ++i) xnorm += H[(size_t)i * m + j] * H[(size_t)i * m + j];
        xnorm = sqrt(xnorm);
        if (xnorm == 0.) {
            tau[j] = 0.;
        } else {
            double beta = -copysign(hypot(alpha, xnorm), alpha);
            tau[j] = (beta - alpha) / beta;
            double scal = 1. / (alpha - beta);
            for (int i = j + 1; i < n; ++i) H[(size_t)i * m + j] *= scal;
            H[(size_t)j * m + j] = beta;
        }
        /* apply H_j = I - tau v v^T to the trailing columns (v_j = 1) */
        for (int c = j + 1; c < m; ++c) {
            double w = H[(size_t)j * m + c];
            for (int i = j + 1; i < n; ++i) w += H[(size_t)i * m + j] * H[(size_t)i * m + c];
            w *= tau[j];
            H[(size_t)j * m + c] -= w;
            for (int i = j + 1; i < n; ++i) H[(size_t)i * m + c] -= w * H[(size_t)i * m + j];
        }
    }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) R[(size_t)i * m + j] = j >= i ? H[(size_t)i * m + j] : 0.;
    /* Q = H_0 H_1 ... H_{m-1} applied to the first m columns of the identity (dorg2r) */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) Q[(size_t)i * m + j] = (i == j) ? 1. : 0.;
    for (int j = m - 1; j >= 0; --j) {
        for (int c = j; c < m; ++c) {
            double w = Q[(size_t)j * m + c];
            for (int i = j + 1; i < n; ++i) w += H[(size_t)i * m + j] * Q[(size_t)i * m + c];
            w *= tau[j];
            Q[(size_t)j * m + c] -= w;
            for (int i = j + 1; i < n; ++i) Q[(size_t)i * m + c] -= w * H[(size_t)i * m + j];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * qgs/toolbox/lyapunov.py:555-632  _compute_backward_lyap_jit / _compute_backward_lyap_traj_jit
 * qgs/toolbox/lyapunov.py:471-552  _compute_forward_lyap_jit  / _compute_forward_lyap_traj_jit
 *
 * One Benettin macro step: propagate the identity over `sub` micro-steps (the reference builds
 * subtime = concatenate(arange(tt, tt+dt, mdt), [tt+dt]) per step; the caller flattens these into
 * sub_time[sub_ptr[i] .. sub_ptr[i+1]) so numpy's arange rounding is reproduced exactly), then
 * q_new = prop @ q, (q, r) = qr(q_new).  The nonlinear state is RESET to the stored trajectory
 * point after each macro step (:601, :622).  q0 (N, n, m) replaces the reference's unseeded
 * np.random.random((n_dim, n_vec)) start (:592-593) -- the caller passes qr(random)[0].
 *
 * `seq` lists, for every Benettin step in execution order, the index of the stored trajectory point
 * to start from (y_idx), the point to reset to afterwards, and the micro-time slice.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    long n_pre;            /* steps before recording starts */
    long n_rec_steps;      /* steps during which records are written */
    const long *start_idx; /* (n_pre + n_rec_steps) index into the stored trajectory used as y */
    const long *sub_ptr;   /* (n_pre + n_rec_steps + 1) */
    const double *sub_time;/* micro times, already directed */
    const double *dt;      /* (n_pre + n_rec_steps) macro dt (signed) used in log|r_ii| / dt */
    long final_idx;        /* stored-trajectory index written into the last record */
} qgso_lyap_plan;

static void benettin_step(const qgso_tensor *T, const double *y, const double *subtime, long nsub,
                          int s, const double *a, const double *b, int adjoint, double inverse,
                          int m, double *q /* (n,m) in/out */, double *r /* (m,m) out */,
                          double *prop, const double *Id, double *qn, double *buf, double *qrwork)
{
    const int n = T->ndim;
    /* write_steps = 0 propagation of the identity: integrate.py:555-614 with tg_ic = Id */
    tgls_one(T, subtime, nsub, y, n, Id, 0, s, a, b, adjoint, inverse, 1, 0, NULL, NULL, NULL, prop, buf);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < m; ++c) {
            double acc = 0.;
            for (int j = 0; j < n; ++j) acc += prop[(size_t)i * n + j] * q[(size_t)j * m + c];
            qn[(size_t)i * m + c] = acc;
        }
    qgso_qr(n, m, qn, q, r, qrwork);
}

/*
 * forward == 0: BLV estimate between tw and t (lyapunov.py:564-632)
 * forward == 1: FLV estimate between t0 and tw (lyapunov.py:480-552)
 * ttraj (N, n, Ltot) is the stored write_steps=1 trajectory (lyapunov.py:558 / :474), computed by
 * the caller with qgso_rk_integrate on concatenate(pretime[:-1], time).
 */
typedef struct {
    const qgso_tensor *T; long Ltot; const double *ttraj; const qgso_lyap_plan *P; int forward, m;
    const double *q0, *r0; long ws; int s; const double *a, *b; int adjoint; double inverse; long R;
    double *rec_traj, *rec_exp, *rec_vec;
} lyap_ctx;

static void lyap_range(void *p, long b0, long e0)
{
    lyap_ctx *c = (lyap_ctx *)p;
    const qgso_tensor *T = c->T;
    const qgso_lyap_plan *P = c->P;
    const int n = T->ndim, m = c->m, s = c->s, forward = c->forward;
    const long Ltot = c->Ltot, R = c->R, ws = c->ws;
    const size_t nm = (size_t)n * m;
    double *buf = (double *)malloc(sizeof(double) * tgls_buf_doubles(n, n, s));
    double *Id = (double *)calloc((size_t)n * n, sizeof(double));
    double *prop = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *q = (double *)malloc(sizeof(double) * nm);
    double *qn = (double *)malloc(sizeof(double) * nm);
    double *r = (double *)calloc((size_t)m * m, sizeof(double));
    double *qrwork = (double *)malloc(sizeof(double) * (nm + m));
    double *y = (double *)malloc(sizeof(double) * n);
    double *mexp = (double *)calloc(m > n ? m : n, sizeof(double));
    for (int i = 0; i < n; ++i) Id[(size_t)i * n + i] = 1.;
    for (long it = b0; it < e0; ++it) {
        const double *tt = c->ttraj + (size_t)it * n * Ltot;
        double *rt = c->rec_traj + (size_t)it * n * R, *re = c->rec_exp + (size_t)it * m * R,
               *rv = c->rec_vec + (size_t)it * nm * R;
        memcpy(q, c->q0 + (size_t)it * nm, sizeof(double) * nm);
        if (c->r0) memcpy(r, c->r0 + (size_t)it * m * m, sizeof(double) * (size_t)m * m);
        else memset(r, 0, sizeof(double) * (size_t)m * m);
        memset(mexp, 0, sizeof(double) * (m > n ? m : n));
        long step = 0;
        for (; step < P->n_pre; ++step) {
            for (int d = 0; d < n; ++d) y[d] = tt[(size_t)d * Ltot + P->start_idx[step]];
            benettin_step(T, y, P->sub_time + P->sub_ptr[step], P->sub_ptr[step + 1] - P->sub_ptr[step],
                          s, c->a, c->b, c->adjoint, c->inverse, m, q, r, prop, Id, qn, buf, qrwork);
        }
        long iw = 0;
        for (long ti = 0; ti < P->n_rec_steps; ++ti, ++step) {
            const long yi = P->start_idx[step];
            for (int cc = 0; cc < m; ++cc)                                   /* :611 / :531 */
                mexp[cc] = log(fabs(r[(size_t)cc * m + cc])) / P->dt[step];
            if (ws > 0 && ti % ws == 0) {
                long col = forward ? R - 1 - iw : iw;                        /* iw -= 1 / iw += 1 */
                for (int cc = 0; cc < m; ++cc) re[(size_t)cc * R + col] = mexp[cc];
                for (int d = 0; d < n; ++d) rt[(size_t)d * R + col] = tt[(size_t)d * Ltot + yi];
                for (size_t e = 0; e < nm; ++e) rv[e * R + col] = q[e];
                ++iw;
            }
            for (int d = 0; d < n; ++d) y[d] = tt[(size_t)d * Ltot + yi];
            benettin_step(T, y, P->sub_time + P->sub_ptr[step], P->sub_ptr[step + 1] - P->sub_ptr[step],
                          s, c->a, c->b, c->adjoint, c->inverse, m, q, r, prop, Id, qn, buf, qrwork);
        }
        {
            /* final record: exponents of the last loop pass, current q (:628-630 / :548-550).
             * Backward: y was reset to traj[ti+1], the point after the last step (:622).
             * Forward: y[0] still holds the point the last pass started from (:527). */
            long col = forward ? 0 : R - 1;
            for (int cc = 0; cc < m; ++cc) re[(size_t)cc * R + col] = mexp[cc];
            for (int d = 0; d < n; ++d) rt[(size_t)d * R + col] = tt[(size_t)d * Ltot + P->final_idx];
            for (size_t e = 0; e < nm; ++e) rv[e * R + col] = q[e];
        }
    }
    free(buf); free(Id); free(prop); free(q); free(qn); free(r); free(qrwork); free(y); free(mexp);
}

int qgso_lyap_benettin(const qgso_tensor *T, long N, long Ltot, const double *ttraj,
                       const qgso_lyap_plan *P, int forward, int m, const double *q0 /* (N,n,m) */,
                       const double *r0 /* (N,m,m) R factor belonging to q0, or NULL */, long ws,
                       int s, const double *a, const double *b, int adjoint, double inverse,
                       long R, double *rec_traj /* (N,n,R) */, double *rec_exp /* (N,m,R) */,
                       double *rec_vec /* (N,n,m,R) */)
{
    const int n = T->ndim;
    memset(rec_traj, 0, sizeof(double) * (size_t)N * n * R);
    memset(rec_exp, 0, sizeof(double) * (size_t)N * m * R);
    memset(rec_vec, 0, sizeof(double) * (size_t)N * n * m * R);
    lyap_ctx cx = {T, Ltot, ttraj, P, forward, m, q0, r0, ws, s, a, b, adjoint, inverse, R,
                   rec_traj, rec_exp, rec_vec};
    parallel_for(N, lyap_range, &cx);
    return 0;
}

int qgso_num_threads(void)
{
    return g_threads > 0 ? g_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
}

void qgso_set_num_threads(int n)
{
    g_threads = n;
}
