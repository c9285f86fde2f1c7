// tgls_reg.cu -- register-resident tangent-linear / Benettin kernels for small bases.
//
// Same arithmetic as tgls.cu (integrate.py:555-614, lyapunov.py:471-632) but laid out for the FP64
// pipe instead of for shared memory: one 64-thread block owns one member, thread c owns COLUMN c of
// the n x m tangent matrix in registers (ndim is a template parameter so every register index is a
// literal).  Per Runge-Kutta stage the block builds the dense Jacobian J(y_s) once in shared memory
// (from the sparse position list) and every column thread computes km[:, c] = +-J @ kms[:, c] with
// broadcast shared-memory reads of J (one LDS.128 feeds two DFMAs of all 32 lanes).  The Benettin
// re-orthonormalisation is a Householder QR (LAPACK dgeqr2 / dorg2r conventions, as np.linalg.qr)
// on the register columns: the owner thread of column j forms the reflector, broadcasts it through
// a double-buffered shared vector (one barrier per reflector), all later columns apply it.
//
// Only "chain" tableaux (a_ij != 0 only for j = i-1; classic RK4, Heun, midpoint, Euler) take this
// path; anything else -- and any ndim without an instantiation -- runs the generic kernels of tgls.cu.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "kernels.cuh"
#include "tgls_shared.cuh"

namespace qgsb {

constexpr int RT = 64;  // threads per block = maximum number of tangent columns

template <int N>
struct RegShared {
    double *xs, *y, *Y, *kst, *yacc, *Jd, *fm, *facc, *vbuf, *rdiag;
    int ldc;  // column stride of fm / facc rows
};

template <int N>
__device__ __forceinline__ RegShared<N> reg_carve(unsigned char *raw, int m)
{
    RegShared<N> S;
    double *p = reinterpret_cast<double *>(raw);
    S.ldc = m;
    S.Jd = p;            p += N * N;          // 16-byte aligned (N even)
    S.xs = p;            p += N + 2;
    S.y = p;             p += N;
    S.Y = p;             p += N;
    S.kst = p;           p += N;
    S.yacc = p;          p += N;
    S.vbuf = p;          p += 2 * (N + 2);
    S.rdiag = p;         p += RT;
    S.fm = p;            p += (size_t)N * m;
    S.facc = p;
    return S;
}

template <int N>
static size_t reg_smem_bytes(int m)
{
    return sizeof(double) * ((size_t)N * N + (N + 2) + 4 * N + 2 * (N + 2) + RT + 2 * (size_t)N * m);
}

// one step of the coupled system for a chain tableau; column registers col[] hold fm[:, c] on entry
// and on exit.  alpha[i] = a[i][i-1].
template <int N, bool ADJ>
__device__ __forceinline__ void reg_tangent_step(const TensorView &T, const TgParams &P, const RegShared<N> &S,
                                                 double dt, double (&col)[N], bool active)
{
    const int tid = threadIdx.x, s = P.s;
    const JacView &J = T.jac;
    double km[N];
    for (int st = 0; st < s; ++st) {
        const double wa_in = st > 0 ? dt * P.a[st * s + st - 1] : 0.;       // (dt a[st]) @ k   integrate.py:216
        const double wb = dt * P.b[st];
        const double wa_out = st + 1 < s ? dt * P.a[(st + 1) * s + st] : 0.;
        if (tid < N) S.xs[tid + 1] = st > 0 ? S.y[tid] + wa_in * S.kst[tid] : S.y[tid];
        __syncthreads();
        if (tid < N) {
            const double k = f_row_rt(T, tid + 1, S.xs);
            S.kst[tid] = k;
            S.yacc[tid] = st == 0 ? wb * k : S.yacc[tid] + wb * k;
        }
        for (int p = tid; p < J.npos; p += RT)
            S.Jd[(J.pos_i[p] - 1) * N + (J.pos_j[p] - 1)] = jac_pos_rt(J, T.rank, p, S.xs);
        __syncthreads();
        if (active) {
            // km = inverse * (J or J^T) @ col        integrate.py:601-603, boundary == 0
#pragma unroll
            for (int i = 0; i < N; ++i) km[i] = 0.;
            if (!ADJ) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double2 *row = reinterpret_cast<const double2 *>(S.Jd + i * N);
#pragma unroll
                    for (int j = 0; j < N / 2; ++j) {
                        const double2 v = row[j];
                        km[i] = fma(v.x, col[2 * j], km[i]);
                        km[i] = fma(v.y, col[2 * j + 1], km[i]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double2 *row = reinterpret_cast<const double2 *>(S.Jd + j * N);
#pragma unroll
                    for (int i = 0; i < N / 2; ++i) {
                        const double2 v = row[i];
                        km[2 * i] = fma(v.x, col[j], km[2 * i]);
                        km[2 * i + 1] = fma(v.y, col[j], km[2 * i + 1]);
                    }
                }
            }
            double *fc = S.facc + tid;
            const double *fmc = S.fm + tid;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double k = P.inverse * km[i];
                const double f = st == 0 ? wb * k : fc[i * S.ldc] + wb * k;     // fm + sum dt b_j km_j  :605-607
                fc[i * S.ldc] = f;
                col[i] = fmc[i * S.ldc] + wa_out * k;                            // km_s of the next stage :598-600
            }
        }
    }
    if (tid < N) S.y[tid] += S.yacc[tid];
    if (active) {
        double *fmc = S.fm + tid;
        const double *fc = S.facc + tid;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            col[i] = fmc[i * S.ldc] + fc[i * S.ldc];
            fmc[i * S.ldc] = col[i];
        }
    }
    __syncthreads();
}

// one nonlinear step of the macro ("stored") trajectory: S.Y <- RK(S.Y, dt)
template <int N>
__device__ __forceinline__ void reg_nl_step(const TensorView &T, const TgParams &P, const RegShared<N> &S, double dt)
{
    const int tid = threadIdx.x, s = P.s;
    for (int st = 0; st < s; ++st) {
        const double wa_in = st > 0 ? dt * P.a[st * s + st - 1] : 0.;
        const double wb = dt * P.b[st];
        if (tid < N) S.xs[tid + 1] = st > 0 ? S.Y[tid] + wa_in * S.kst[tid] : S.Y[tid];
        __syncthreads();
        if (tid < N) {
            const double k = f_row_rt(T, tid + 1, S.xs);
            S.kst[tid] = k;
            S.yacc[tid] = st == 0 ? wb * k : S.yacc[tid] + wb * k;
        }
        __syncthreads();
    }
    if (tid < N) S.Y[tid] += S.yacc[tid];
    __syncthreads();
}

// Householder QR of the n x m matrix whose column c lives in col[] of thread c.  On exit col[] holds
// Q[:, c]; S.rdiag the diagonal of R; Rout (m x m row-major, may be null) the whole factor.
template <int N>
__device__ __forceinline__ void reg_qr(const RegShared<N> &S, int m, double (&col)[N], bool active, double *Rout)
{
    const int tid = threadIdx.x;
    double q[N];
    double mytau = 0.;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (j < m) {                                   // uniform
            double *buf = S.vbuf + (j & 1) * (N + 2);
            if (tid == j) {
                // dlarfg on col[j..N-1]
                double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
#pragma unroll
                for (int i = j + 1; i < N; ++i) {
                    const double v = col[i] * col[i];
                    if (((i - j) & 3) == 0) s0 += v;
                    else if (((i - j) & 3) == 1) s1 += v;
                    else if (((i - j) & 3) == 2) s2 += v;
                    else s3 += v;
                }
                const double xnorm = sqrt((s0 + s1) + (s2 + s3));
                const double alpha = col[j];
                double beta = alpha, tau = 0.;
                if (xnorm != 0.) {
                    beta = -copysign(hypot(alpha, xnorm), alpha);
                    tau = (beta - alpha) / beta;
                    const double scal = 1. / (alpha - beta);
#pragma unroll
                    for (int i = j + 1; i < N; ++i) col[i] *= scal;
                }
                col[j] = beta;
                mytau = tau;
#pragma unroll
                for (int i = j + 1; i < N; ++i) buf[i] = col[i];
                buf[N] = tau;
                S.rdiag[j] = beta;
            }
            __syncthreads();
            if (active && tid > j) {
                const double tau = buf[N];
                double w = col[j];
#pragma unroll
                for (int i = j + 1; i < N; ++i) w = fma(buf[i], col[i], w);
                w *= tau;
                col[j] -= w;
#pragma unroll
                for (int i = j + 1; i < N; ++i) col[i] = fma(-w, buf[i], col[i]);
            }
            if (Rout != nullptr && active && tid >= j) Rout[j * m + tid] = col[j];
        }
    }
    if (Rout != nullptr && active) {
        for (int i = tid + 1; i < m; ++i) Rout[i * m + tid] = 0.;   // strictly lower part of column tid
    }
    // dorg2r: Q = H_0 ... H_{m-1} [I; 0]
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = i == tid ? 1. : 0.;
#pragma unroll
    for (int jj = 0; jj < N; ++jj) {
        const int j = N - 1 - jj;
        if (j < m) {
            double *buf = S.vbuf + (j & 1) * (N + 2);
            if (tid == j) {
#pragma unroll
                for (int i = j + 1; i < N; ++i) buf[i] = col[i];
                buf[N] = mytau;
            }
            __syncthreads();
            if (active && tid >= j) {
                const double tau = buf[N];
                double w = q[j];
#pragma unroll
                for (int i = j + 1; i < N; ++i) w = fma(buf[i], q[i], w);
                w *= tau;
                q[j] -= w;
#pragma unroll
                for (int i = j + 1; i < N; ++i) q[i] = fma(-w, buf[i], q[i]);
            }
        }
    }
    __syncthreads();
    if (active) {
        double *fmc = S.fm + tid;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            col[i] = q[i];
            fmc[i * S.ldc] = q[i];
        }
    }
    __syncthreads();
}

template <int N>
__device__ __forceinline__ void reg_load_common(const TensorView &T, const RegShared<N> &S)
{
    for (int q = threadIdx.x; q < N * N; q += RT) S.Jd[q] = 0.;    // structural zeros stay zero
    if (threadIdx.x < N) {
        S.kst[threadIdx.x] = 0.;
        S.yacc[threadIdx.x] = 0.;
    }
    if (threadIdx.x == 0) S.xs[0] = 1.;
}

// ---- plain tangent-linear integration (integrate.py:555-614) ------------------------------------------------
template <int N, bool ADJ>
__global__ void __launch_bounds__(RT) tgls_reg_kernel(TensorView T, const __grid_constant__ TgParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long member = blockIdx.x;
    const int m = P.m, tid = threadIdx.x, nm = N * m;
    const bool active = tid < m;
    RegShared<N> S = reg_carve<N>(smem_raw, m);
    double col[N];
    reg_load_common<N>(T, S);
    if (tid < N) S.y[tid] = P.y[member * N + tid];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        col[i] = active ? P.fm[member * nm + i * m + tid] : 0.;
        if (active) S.fm[i * S.ldc + tid] = col[i];
    }
    __syncthreads();
    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        if (P.rec_y && P.write_steps > 0 && ti % P.write_steps == 0) {
            double *ry = P.rec_y + ((size_t)iw * P.n_members + member) * N;
            double *rf = P.rec_fm + ((size_t)iw * P.n_members + member) * nm;
            if (tid < N) ry[tid] = S.y[tid];
            if (active) {
#pragma unroll
                for (int i = 0; i < N; ++i) rf[i * m + tid] = col[i];
            }
            ++iw;
        }
        reg_tangent_step<N, ADJ>(T, P, S, P.dt[ti], col, active);
    }
    if (P.rec_y) {
        double *ry = P.rec_y + ((size_t)(P.n_records - 1) * P.n_members + member) * N;
        double *rf = P.rec_fm + ((size_t)(P.n_records - 1) * P.n_members + member) * nm;
        if (tid < N) ry[tid] = S.y[tid];
        if (active) {
#pragma unroll
            for (int i = 0; i < N; ++i) rf[i * m + tid] = col[i];
        }
    }
    if (tid < N) P.y[member * N + tid] = S.y[tid];
    if (active) {
#pragma unroll
        for (int i = 0; i < N; ++i) P.fm[member * nm + i * m + tid] = col[i];
    }
}

// ---- Benettin loop (lyapunov.py:471-632) -----------------------------------------------------------------------
template <int N, bool ADJ>
__global__ void __launch_bounds__(RT) lyap_reg_kernel(TensorView T, const __grid_constant__ TgParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long member = blockIdx.x;
    const int m = P.m, tid = threadIdx.x, nm = N * m;
    const bool active = tid < m;
    RegShared<N> S = reg_carve<N>(smem_raw, m);
    const long steps = P.n_pre + P.n_rec;
    const long R = P.n_records;
    double col[N];
    reg_load_common<N>(T, S);
    if (tid < N) S.Y[tid] = P.y[member * N + tid];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        col[i] = active ? P.fm[member * nm + i * m + tid] : 0.;
        if (active) S.fm[i * S.ldc + tid] = col[i];
    }
    if (active) S.rdiag[tid] = P.r0 ? P.r0[((size_t)member * m + tid) * m + tid] : 0.;
    __syncthreads();
    const size_t sbase = P.stored ? tile_base(member, N) : 0;
    long iw = 0;
    double mexp = 0.;
    for (long step = 0; step < steps; ++step) {
        if (P.stored) {                                                   // lyapunov.py:513 / :527
            const double *src = P.stored + (size_t)P.start_idx[step] * N * P.stored_ld + sbase;
            if (tid < N) S.Y[tid] = src[(size_t)tid * TILE];
            __syncthreads();
        }
        if (step >= P.n_pre) {
            const long ti = step - P.n_pre;
            if (active) mexp = log(fabs(S.rdiag[tid])) / P.dt_macro[step];   // :611 / :531
            if (P.q_all && active) {
                double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + ti) * nm;
#pragma unroll
                for (int i = 0; i < N; ++i) qa[i * m + tid] = col[i];
            }
            if (P.write_steps > 0 && ti % P.write_steps == 0) {
                const long c = P.forward == 1 ? R - 1 - iw : iw;
                double *ry = P.rec_y + ((size_t)c * P.n_members + member) * N;
                double *rv = P.rec_fm + ((size_t)c * P.n_members + member) * nm;
                double *re = P.rec_exp + ((size_t)c * P.n_members + member) * m;
                if (tid < N) ry[tid] = S.Y[tid];
                if (active) {
                    if (P.rec_fm) {
#pragma unroll
                        for (int i = 0; i < N; ++i) rv[i * m + tid] = col[i];
                    }
                    re[tid] = mexp;
                }
                ++iw;
            }
        }
        // propagate the basis over the micro steps starting from the stored point (:598-600)
        if (tid < N) S.y[tid] = S.Y[tid];
        __syncthreads();
        for (long q = P.sub_ptr[step]; q < P.sub_ptr[step + 1]; ++q)
            reg_tangent_step<N, ADJ>(T, P, S, P.sub_dt[q], col, active);
        // q, r = qr(prop @ q)   (:602-604)
        reg_qr<N>(S, m, col, active, (P.r_all && step >= P.r_first) ? P.r_all + ((size_t)member * (steps - P.r_first) + (step - P.r_first)) * m * m : nullptr);
        if (P.forward == 2) {                             // Ginelli forward pass: follow the micro steps
            if (tid < N) S.Y[tid] = S.y[tid];
            __syncthreads();
        } else if (!P.stored) {                           // next stored-trajectory point (:601 / :622)
            reg_nl_step<N>(T, P, S, P.dt_macro[step]);
        }
    }
    {
        if (P.stored) {
            const double *src = P.stored + (size_t)P.final_idx * N * P.stored_ld + sbase;
            if (tid < N) S.Y[tid] = src[(size_t)tid * TILE];
            __syncthreads();
        }
        const long c = P.forward == 1 ? 0 : R - 1;                        // :628-630 / :548-550
        double *ry = P.rec_y + ((size_t)c * P.n_members + member) * N;
        double *rv = P.rec_fm + ((size_t)c * P.n_members + member) * nm;
        double *re = P.rec_exp + ((size_t)c * P.n_members + member) * m;
        if (tid < N) ry[tid] = S.Y[tid];
        if (active) {
            if (P.rec_fm) {
#pragma unroll
                for (int i = 0; i < N; ++i) rv[i * m + tid] = col[i];
            }
            re[tid] = mexp;
            if (P.q_all) {
                double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + P.n_rec) * nm;
#pragma unroll
                for (int i = 0; i < N; ++i) qa[i * m + tid] = col[i];
            }
        }
    }
    if (tid < N) P.y[member * N + tid] = S.Y[tid];
    if (active) {
#pragma unroll
        for (int i = 0; i < N; ++i) P.fm[member * nm + i * m + tid] = col[i];
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
template <typename K>
static void launch_one(K kernel, const qgsb_tensor *t, const TgParams &P, size_t bytes)
{
    QGSB_REQUIRE(bytes <= ctx().smem_optin, "register tangent kernel needs %zu bytes of shared memory", bytes);
    if (bytes > 48 * 1024)
        QGSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kernel<<<(unsigned)P.n_members, RT, bytes, ctx().stream>>>(t->view, P);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

template <int N>
static void launch_n(const qgsb_tensor *t, const TgParams &P, bool lyap)
{
    const size_t bytes = reg_smem_bytes<N>(P.m);
    const bool adj = P.adjoint != 0;
    if (lyap) {
        if (adj) launch_one(lyap_reg_kernel<N, true>, t, P, bytes);
        else launch_one(lyap_reg_kernel<N, false>, t, P, bytes);
    } else {
        if (adj) launch_one(tgls_reg_kernel<N, true>, t, P, bytes);
        else launch_one(tgls_reg_kernel<N, false>, t, P, bytes);
    }
}

bool reg_tangent_supported(const qgsb_tensor *t, const Tableau &tab, int m)
{
    const char *force = getenv("QGSB_TGLS_GENERIC");
    if (force && force[0] == '1') return false;
    const char *which = getenv("QGSB_TGLS_KERNEL");
    if (which && !strcmp(which, "generic")) return false;
    const int n = t->view.n;
    // measured on B200: below ~16 columns the shared-memory kernel of tgls.cu is faster (the register kernel always
    // pays for the dense n x n product per column thread and for two mostly idle warps)
    return tab.chain && m >= 16 && m <= RT && (n == 20 || n == 36 || n == 38);
}

void launch_reg_tangent(const qgsb_tensor *t, const TgParams &P, bool lyap)
{
    switch (t->view.n) {
        case 20: return launch_n<20>(t, P, lyap);
        case 36: return launch_n<36>(t, P, lyap);
        case 38: return launch_n<38>(t, P, lyap);
        default: QGSB_REQUIRE(false, "no register tangent kernel for ndim %d", t->view.n);
    }
}

}  // namespace qgsb
