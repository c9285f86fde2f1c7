// tgls.cu -- tangent-linear / adjoint propagation and Benettin re-orthonormalisation kernels.
//
// Replaces _integrate_runge_kutta_tgls_jit (qgs/integrators/integrate.py:555-614), the
// TglsTrajectoryProcess pool (qgs/integrators/integrator.py:1103-1169) and the per-member Benettin
// loops of qgs/toolbox/lyapunov.py:471-632 (propagate, np.linalg.qr, log|diag R| / dt).
//
// One thread block owns one ensemble member.  The nonlinear state, the Jacobian (as the values of
// its structurally non-zero positions, never as a dense matrix) and the n x m tangent matrices of
// all Runge-Kutta stages live in shared memory; the product J @ X walks the CSR list of Jacobian
// positions (CSC for the adjoint), so the work is nnz(J) * m instead of the reference's dense n*n*m.
// Data layout in HBM is member-major ((N, n), (N, n, m)), i.e. the API layout: a block reads and
// writes contiguous chunks.
#include <algorithm>
#include <cmath>
#include <memory>

#include "common.cuh"
#include "kernels.cuh"
#include "tgls_shared.cuh"

namespace qgsb {

struct TgShared {
    double *xs, *y, *Y, *K, *Jv, *rdiag, *mexp, *red, *fm, *kms, *KM;
    double *Jd;      // dense Jacobian of the member in the global scratch (large bases), else null
    double *tile;    // shared-memory staging of one 32 x 32 block of the stage input (dense product)
};

constexpr int DP_KB = 32;             // rows of the stage input staged per block
constexpr int DP_NB = 32;             // columns per chunk (four 8-wide tensor-core tiles)
constexpr int DP_STRIDE = 36;         // row stride of the staged block in doubles: the 16 lanes of a half warp read
                                      // rows 0..3 x columns 0..3 of a tile -- 36 = 4 mod 16 spreads them over all banks
constexpr int DP_TILE_DOUBLES = DP_KB * DP_STRIDE;

__device__ __forceinline__ TgShared carve(unsigned char *raw, const TensorView &T, const TgParams &P, long member)
{
    const int n = T.n, m = P.m, s = P.s;
    TgShared S;
    double *p = reinterpret_cast<double *>(raw);
    S.xs = p;            p += n + 1;
    S.y = p;             p += n;
    S.Y = p;             p += n;
    S.K = p;             p += (size_t)s * n;
    S.Jv = p;            p += P.jd_ld ? 0 : T.jac.npos;     // dense J lives in the scratch instead
    S.tile = p;          p += P.jd_ld ? DP_TILE_DOUBLES : 0;
    S.rdiag = p;         p += m;
    S.mexp = p;          p += m;      // local exponent of every vector at the last macro step
    S.red = p;           p += 64;
    double *mat = P.scratch ? P.scratch + (size_t)member * P.scratch_per_member : p;
    S.fm = mat;
    S.kms = mat + (size_t)n * m;
    S.KM = mat + (size_t)2 * n * m;
    S.Jd = P.jd_ld ? mat + (size_t)(s + 2) * n * m : nullptr;
    return S;
}

// one explicit Runge-Kutta step of the nonlinear state only: S.y <- RK(S.y, dt)
template <int RANK>
__device__ void nl_step(const TensorView &T, const TgParams &P, const TgShared &S, double dt)
{
    const int n = T.n, s = P.s, tid = threadIdx.x;
    for (int st = 0; st < s; ++st) {
        for (int r = tid; r < n; r += TG_THREADS) {
            double v = 0.;
            for (int j = 0; j < st; ++j) {
                const double w = P.a[st * s + j];
                if (w != 0.) v += (dt * w) * S.K[(size_t)j * n + r];
            }
            S.xs[r + 1] = S.y[r] + v;
        }
        __syncthreads();
        for (int r = tid; r < n; r += TG_THREADS) S.K[(size_t)st * n + r] = f_row<RANK>(T, r + 1, S.xs);
        __syncthreads();
    }
    for (int r = tid; r < n; r += TG_THREADS) {
        double v = 0.;
        for (int j = 0; j < s; ++j) v += (dt * P.b[j]) * S.K[(size_t)j * n + r];
        S.y[r] += v;
    }
    __syncthreads();
}

// ---- large bases: out = scale * Jd @ X on the FP64 tensor cores ---------------------------------------------------------
// BASELINE.json allows the tensor cores "only where a dense formulation is measured to beat the sparse kernel for large
// bases".  For the TENDENCIES it never does (section 4.2 of DESIGN.md: 100x the flops at the same peak).  For the
// tangent product it does: the Jacobian of the 228-variable model has 26 140 structurally non-zero positions of 51 984
// (50 % dense, 79 % of the 8 x 4 operand tiles non-empty), the sparse product walks them once per COLUMN of X with
// operands in L2, and the dense product is a GEMM, (n x n) @ (n x m) per stage and member, whose operands are reused
// from registers and shared memory.  mma.sync.m8n8k4.f64: A (8 x 4) lane l holds A[l / 4][l % 4]; B (4 x 8) lane l holds
// B[l % 4][l / 4]; C (8 x 8) lane l holds C[l / 4][2 (l % 4) + {0, 1}].
// Jd is (ld x ld) row-major with rows / columns >= n zero; X and out are (n x m) row-major.  A warp owns row tiles
// w, w + 4, ... (at most 8) and sweeps the columns in chunks of 32: the 32 x 32 block of X of every k step is staged in
// shared memory once for the four warps, the A fragments come straight from L2 (each is used for the four column tiles
// of the chunk).
__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ void dense_product(const double *__restrict__ Jd, int ld, const double *__restrict__ X, double *__restrict__ out,
                              int n, int m, double scale, double *tile)
{
    constexpr int WARPS = TG_THREADS / 32, MAX_RT = 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_tiles = ld / 8;
    const int ar = lane >> 2, ac = lane & 3;          // A: row in tile, column in k chunk;  B: ac = row in k chunk, ar = column
    for (int c0 = 0; c0 < m; c0 += DP_NB) {
        double acc[MAX_RT][4][2];
#pragma unroll
        for (int r = 0; r < MAX_RT; ++r)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[r][t][0] = acc[r][t][1] = 0.;
        for (int kb = 0; kb < ld; kb += DP_KB) {
            __syncthreads();                           // the previous block has been consumed
            for (int q = tid; q < DP_KB * DP_NB; q += TG_THREADS) {
                const int kr = q / DP_NB, cc = q - kr * DP_NB;
                const int row = kb + kr, col = c0 + cc;
                tile[kr * DP_STRIDE + cc] = (row < n && col < m) ? X[(size_t)row * m + col] : 0.;
            }
            __syncthreads();
#pragma unroll 2
            for (int kk = 0; kk < DP_KB / 4; ++kk) {
                if (kb + kk * 4 >= ld) break;
                double bf[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) bf[t] = tile[(kk * 4 + ac) * DP_STRIDE + t * 8 + ar];
                double af[MAX_RT];
#pragma unroll
                for (int r = 0; r < MAX_RT; ++r) {
                    const int rt = warp + r * WARPS;
                    af[r] = rt < row_tiles ? Jd[(size_t)(rt * 8 + ar) * ld + kb + kk * 4 + ac] : 0.;
                }
#pragma unroll
                for (int r = 0; r < MAX_RT; ++r)
#pragma unroll
                    for (int t = 0; t < 4; ++t) dmma_884(acc[r][t][0], acc[r][t][1], af[r], bf[t]);
            }
        }
#pragma unroll
        for (int r = 0; r < MAX_RT; ++r) {
            const int row = (warp + r * WARPS) * 8 + ar;
            if (row >= n) continue;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int col = c0 + t * 8 + 2 * ac;
                if (col < m) out[(size_t)row * m + col] = scale * acc[r][t][0];
                if (col + 1 < m) out[(size_t)row * m + col + 1] = scale * acc[r][t][1];
            }
        }
    }
    __syncthreads();
}

// one step of the coupled system (integrate.py:590-609): S.y and S.fm advance by dt
template <int RANK, bool DENSE>
__device__ void tg_step(const TensorView &T, const TgParams &P, const TgShared &S, double dt)
{
    const int n = T.n, m = P.m, s = P.s, tid = threadIdx.x;
    const int nm = n * m;
    const JacView &J = T.jac;
    for (int st = 0; st < s; ++st) {
        // stage state of the nonlinear part and of the tangent part
        for (int r = tid; r < n; r += TG_THREADS) {
            double v = 0.;
            for (int j = 0; j < st; ++j) {
                const double w = P.a[st * s + j];
                if (w != 0.) v += (dt * w) * S.K[(size_t)j * n + r];
            }
            S.xs[r + 1] = S.y[r] + v;
        }
        for (int q = tid; q < nm; q += TG_THREADS) {
            double v = S.fm[q];                                            // km_s = fm + sum dt a_ij km_j  :598-600
            for (int j = 0; j < st; ++j) {
                const double w = P.a[st * s + j];
                if (w != 0.) v += (dt * w) * S.KM[(size_t)j * nm + q];
            }
            S.kms[q] = v;
        }
        __syncthreads();
        // tendencies and Jacobian positions at the stage state
        for (int r = tid; r < n; r += TG_THREADS) S.K[(size_t)st * n + r] = f_row<RANK>(T, r + 1, S.xs);
        double *out = S.KM + (size_t)st * nm;
        if (DENSE) {
            // large bases: the position values go into the member's dense matrix (transposed for the adjoint; the
            // structural zeros were written once at kernel start) and the product runs on the tensor cores
            const int ld = P.jd_ld;
            for (int p = tid; p < J.npos; p += TG_THREADS) {
                const int i = J.pos_i[p] - 1, j = J.pos_j[p] - 1;
                S.Jd[P.adjoint ? (size_t)j * ld + i : (size_t)i * ld + j] = jac_pos<RANK>(J, p, S.xs);
            }
            __syncthreads();
            dense_product(S.Jd, ld, S.kms, out, n, m, P.inverse, S.tile);
            continue;
        }
        for (int p = tid; p < J.npos; p += TG_THREADS) S.Jv[p] = jac_pos<RANK>(J, p, S.xs);
        __syncthreads();
        // km_i = inverse * (J or J^T) @ km_s          :601-603 with boundary == 0
        for (int q = tid; q < nm; q += TG_THREADS) {
            const int r = q / m + 1, c = q - (r - 1) * m;
            double acc = 0.;
            if (!P.adjoint) {
                for (int p = J.row_ptr[r]; p < J.row_ptr[r + 1]; ++p)
                    acc += S.Jv[p] * S.kms[(size_t)(J.pos_j[p] - 1) * m + c];
            } else {
                for (int pp = J.col_ptr[r]; pp < J.col_ptr[r + 1]; ++pp) {
                    const int p = J.col_perm[pp];
                    acc += S.Jv[p] * S.kms[(size_t)(J.pos_i[p] - 1) * m + c];
                }
            }
            out[q] = P.inverse * acc;
        }
        __syncthreads();
    }
    for (int r = tid; r < n; r += TG_THREADS) {
        double v = 0.;
        for (int j = 0; j < s; ++j) v += (dt * P.b[j]) * S.K[(size_t)j * n + r];
        S.y[r] += v;
    }
    for (int q = tid; q < nm; q += TG_THREADS) {
        double v = S.fm[q];
        for (int j = 0; j < s; ++j) v += (dt * P.b[j]) * S.KM[(size_t)j * nm + q];   // :605-607
        S.fm[q] = v;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// plain TGLS integration
// ------------------------------------------------------------------------------------------------
template <int RANK, bool DENSE>
__global__ void __launch_bounds__(TG_THREADS) tgls_kernel(TensorView T, const __grid_constant__ TgParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long member = blockIdx.x;
    const int n = T.n, m = P.m, tid = threadIdx.x, nm = n * m;
    TgShared S = carve(smem_raw, T, P, member);
    for (int r = tid; r < n; r += TG_THREADS) S.y[r] = P.y[member * n + r];
    for (int q = tid; q < nm; q += TG_THREADS) S.fm[q] = P.fm[member * nm + q];
    if (DENSE)
        for (int q = tid; q < P.jd_ld * P.jd_ld; q += TG_THREADS) S.Jd[q] = 0.;     // structural zeros, written once
    if (tid == 0) S.xs[0] = 1.;
    __syncthreads();
    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        if (P.rec_y && P.write_steps > 0 && ti % P.write_steps == 0) {      // integrate.py:585-588
            double *ry = P.rec_y + ((size_t)iw * P.n_members + member) * n;
            double *rf = P.rec_fm + ((size_t)iw * P.n_members + member) * nm;
            for (int r = tid; r < n; r += TG_THREADS) ry[r] = S.y[r];
            for (int q = tid; q < nm; q += TG_THREADS) rf[q] = S.fm[q];
            ++iw;
        }
        tg_step<RANK, DENSE>(T, P, S, P.dt[ti]);
    }
    if (P.rec_y) {                                                           // integrate.py:611-612
        double *ry = P.rec_y + ((size_t)(P.n_records - 1) * P.n_members + member) * n;
        double *rf = P.rec_fm + ((size_t)(P.n_records - 1) * P.n_members + member) * nm;
        for (int r = tid; r < n; r += TG_THREADS) ry[r] = S.y[r];
        for (int q = tid; q < nm; q += TG_THREADS) rf[q] = S.fm[q];
    }
    for (int r = tid; r < n; r += TG_THREADS) P.y[member * n + r] = S.y[r];
    for (int q = tid; q < nm; q += TG_THREADS) P.fm[member * nm + q] = S.fm[q];
}

// ------------------------------------------------------------------------------------------------
// Householder QR of the n x m matrix A (row-major, in place) with LAPACK's dgeqr2 / dorg2r
// conventions -- np.linalg.qr of lyapunov.py:602-604.  On exit A holds Q (n x m), Rout (m x m,
// row-major, may be null) the upper-triangular factor and rdiag its diagonal.  W is a workspace of
// n m + 2 m doubles, red three doubles of scratch.
// ------------------------------------------------------------------------------------------------
__device__ void block_qr(int n, int m, double *A, double *W, double *rdiag, double *red, double *Rout)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // workspace carving: tau_j in W[0..m), column weights in W[m..2m), Q built in W[2m..2m + n m).
    // (W is the stage-input buffer followed by the stage-derivative buffers, all free during the QR.)
    double *taus = W;
    double *wv = W + m;
    // column dot products use 4 threads per column (rows strided by 4) and two shuffle steps
    const int part = tid & 3, cslot = tid >> 2;           // TG_THREADS / 4 = 32 column slots per pass
    for (int j = 0; j < m; ++j) {
        // ---- reflector for column j (dgeqr2 / dlarfg) ----
        if (warp == 0) {
            double ss = 0.;
            for (int i = j + 1 + lane; i < n; i += 32) {
                const double v = A[(size_t)i * m + j];
                ss += v * v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) {
                const double alpha = A[(size_t)j * m + j];
                const double xnorm = sqrt(ss);
                if (xnorm == 0.) {
                    red[0] = 0.;        // tau
                    red[1] = 0.;        // 1 / (alpha - beta)
                    red[2] = alpha;     // beta
                } else {
                    const double beta = -copysign(hypot(alpha, xnorm), alpha);
                    red[0] = (beta - alpha) / beta;
                    red[1] = 1. / (alpha - beta);
                    red[2] = beta;
                }
            }
        }
        __syncthreads();
        const double tj = red[0], scal = red[1], beta = red[2];
        // ---- w_c = tau * (A[j][c] + scal * sum_{i>j} x_i A[i][c]) for the trailing columns (v = [1; scal x]) ----
        for (int c0 = j + 1; c0 < m; c0 += TG_THREADS / 4) {
            const int c = c0 + cslot;
            double w = 0.;
            if (c < m)
                for (int i = j + 1 + part; i < n; i += 4) w += A[(size_t)i * m + j] * A[(size_t)i * m + c];
            w += __shfl_xor_sync(0xffffffffu, w, 1);
            w += __shfl_xor_sync(0xffffffffu, w, 2);
            if (c < m && part == 0) wv[c] = tj * (A[(size_t)j * m + c] + scal * w);
        }
        __syncthreads();
        // ---- A[:, c] -= w_c v for c > j; store v (scaled) and beta in column j ----
        const int rows = n - j, cols = m - j;      // column slot 0 is column j itself
        for (int q = tid; q < rows * cols; q += TG_THREADS) {
            const int i = j + q / cols, c = j + q % cols;
            if (c == j) {
                if (i == j) {
                    A[(size_t)j * m + j] = beta;
                    rdiag[j] = beta;
                    taus[j] = tj;
                }
            } else {
                const double vi = i == j ? 1. : A[(size_t)i * m + j] * scal;
                A[(size_t)i * m + c] -= wv[c] * vi;
            }
        }
        __syncthreads();
        // scaling of column j happens after every reader of the unscaled x is done
        if (tj != 0.)
            for (int i = j + 1 + tid; i < n; i += TG_THREADS) A[(size_t)i * m + j] *= scal;
        // (the next iteration's first barrier orders this against later readers of column j: none before Q)
    }
    __syncthreads();
    if (Rout)
        for (int q = tid; q < m * m; q += TG_THREADS) {
            const int i = q / m, c = q % m;
            Rout[q] = c >= i ? A[(size_t)i * m + c] : 0.;
        }
    // ---- form Q = H_0 ... H_{m-1} [I; 0]   (dorg2r) in W[2m..), then copy back ----
    double *Q = W + 2 * (size_t)m;
    for (int q = tid; q < n * m; q += TG_THREADS) Q[q] = (q / m == q % m) ? 1. : 0.;
    __syncthreads();
    for (int j = m - 1; j >= 0; --j) {
        const double tj = taus[j];
        // columns c < j of Q are still those of the identity below row j: only c >= j change
        for (int c0 = j; c0 < m; c0 += TG_THREADS / 4) {
            const int c = c0 + cslot;
            double w = 0.;
            if (c < m)
                for (int i = j + 1 + part; i < n; i += 4) w += A[(size_t)i * m + j] * Q[(size_t)i * m + c];
            w += __shfl_xor_sync(0xffffffffu, w, 1);
            w += __shfl_xor_sync(0xffffffffu, w, 2);
            if (c < m && part == 0) wv[c] = tj * (Q[(size_t)j * m + c] + w);
        }
        __syncthreads();
        const int rows = n - j, cols = m - j;
        for (int q = tid; q < rows * cols; q += TG_THREADS) {
            const int i = j + q / cols, c = j + q % cols;
            const double vi = i == j ? 1. : A[(size_t)i * m + j];
            Q[(size_t)i * m + c] -= wv[c] * vi;
        }
        __syncthreads();
    }
    for (int q = tid; q < n * m; q += TG_THREADS) A[q] = Q[q];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Benettin loop
// ------------------------------------------------------------------------------------------------
template <int RANK, bool DENSE>
__global__ void __launch_bounds__(TG_THREADS) lyap_kernel(TensorView T, const __grid_constant__ TgParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long member = blockIdx.x;
    const int n = T.n, m = P.m, tid = threadIdx.x, nm = n * m;
    TgShared S = carve(smem_raw, T, P, member);
    const long steps = P.n_pre + P.n_rec;
    const long R = P.n_records;
    // Y: macro (stored-trajectory) state; y: micro state of the tangent propagation
    for (int r = tid; r < n; r += TG_THREADS) S.Y[r] = P.y[member * n + r];
    for (int q = tid; q < nm; q += TG_THREADS) S.fm[q] = P.fm[member * nm + q];
    for (int c = tid; c < m; c += TG_THREADS) S.rdiag[c] = P.r0 ? P.r0[((size_t)member * m + c) * m + c] : 0.;
    if (DENSE)
        for (int q = tid; q < P.jd_ld * P.jd_ld; q += TG_THREADS) S.Jd[q] = 0.;     // structural zeros, written once
    if (tid == 0) S.xs[0] = 1.;
    __syncthreads();
    const size_t sbase = P.stored ? tile_base(member, n) : 0;
    long iw = 0;
    for (int c = tid; c < m; c += TG_THREADS) S.mexp[c] = 0.;
    // step -1 (P.qr_at_start): only factorise the start matrix drawn on the device (lyapunov.py:592-593)
    for (long step = P.qr_at_start ? -1 : 0; step < steps; ++step) {
        const bool real = step >= 0;
        if (real && P.stored) {                                           // lyapunov.py:513 / :527
            const double *src = P.stored + (size_t)P.start_idx[step] * n * P.stored_ld + sbase;
            for (int r = tid; r < n; r += TG_THREADS) S.Y[r] = src[(size_t)r * TILE];
            __syncthreads();
        }
        if (real && step >= P.n_pre) {
            const long ti = step - P.n_pre;
            for (int c = tid; c < m; c += TG_THREADS) S.mexp[c] = log(fabs(S.rdiag[c])) / P.dt_macro[step];   // :611 / :531
            if (P.q_all) {
                double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + ti) * nm;
                for (int q = tid; q < nm; q += TG_THREADS) qa[q] = S.fm[q];
            }
            if (P.write_steps > 0 && ti % P.write_steps == 0) {
                const long col = P.forward == 1 ? R - 1 - iw : iw;
                double *ry = P.rec_y + ((size_t)col * P.n_members + member) * n;
                double *rv = P.rec_fm + ((size_t)col * P.n_members + member) * nm;
                double *re = P.rec_exp + ((size_t)col * P.n_members + member) * m;
                for (int r = tid; r < n; r += TG_THREADS) ry[r] = S.Y[r];
                if (P.rec_fm)
                    for (int q = tid; q < nm; q += TG_THREADS) rv[q] = S.fm[q];
                for (int c = tid; c < m; c += TG_THREADS) re[c] = S.mexp[c];
                ++iw;
            }
        }
        // propagate the basis over the micro steps starting from the stored point (:598-600)
        if (real) {
            for (int r = tid; r < n; r += TG_THREADS) S.y[r] = S.Y[r];
            __syncthreads();
            for (long q = P.sub_ptr[step]; q < P.sub_ptr[step + 1]; ++q) tg_step<RANK, DENSE>(T, P, S, P.sub_dt[q]);
        }
        // q_new = prop @ q ; q, r = qr(q_new)   (:602-604) -- fm already holds prop @ q by linearity
        block_qr(n, m, S.fm, S.kms, S.rdiag, S.red, (real && P.r_all && step >= P.r_first) ? P.r_all + ((size_t)member * (steps - P.r_first) + (step - P.r_first)) * m * m : nullptr);
        if (!real) continue;
        // next point of the stored trajectory (:601 / :622): one nonlinear step of length dt_macro
        if (P.forward == 2) {
            // Ginelli forward pass (lyapunov.py:1212-1218): the trajectory follows the micro-steps
            for (int r = tid; r < n; r += TG_THREADS) S.Y[r] = S.y[r];
            __syncthreads();
        } else if (!P.stored) {
            for (int r = tid; r < n; r += TG_THREADS) S.y[r] = S.Y[r];
            __syncthreads();
            nl_step<RANK>(T, P, S, P.dt_macro[step]);
            for (int r = tid; r < n; r += TG_THREADS) S.Y[r] = S.y[r];
            __syncthreads();
        }
    }
    {
        // final record (:628-630 / :548-550)
        if (P.stored) {
            const double *src = P.stored + (size_t)P.final_idx * n * P.stored_ld + sbase;
            for (int r = tid; r < n; r += TG_THREADS) S.Y[r] = src[(size_t)r * TILE];
            __syncthreads();
        }
        const long col = P.forward == 1 ? 0 : R - 1;
        double *ry = P.rec_y + ((size_t)col * P.n_members + member) * n;
        double *rv = P.rec_fm + ((size_t)col * P.n_members + member) * nm;
        double *re = P.rec_exp + ((size_t)col * P.n_members + member) * m;
        for (int r = tid; r < n; r += TG_THREADS) ry[r] = S.Y[r];
        if (P.rec_fm)
            for (int q = tid; q < nm; q += TG_THREADS) rv[q] = S.fm[q];
        for (int c = tid; c < m; c += TG_THREADS) re[c] = S.mexp[c];
        if (P.q_all) {
            double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + P.n_rec) * nm;
            for (int q = tid; q < nm; q += TG_THREADS) qa[q] = S.fm[q];
        }
    }
    for (int r = tid; r < n; r += TG_THREADS) P.y[member * n + r] = S.Y[r];
    for (int q = tid; q < nm; q += TG_THREADS) P.fm[member * nm + q] = S.fm[q];
}

// (R, X) -> (X, R) with optional flip of the record axis
__global__ void transpose_rec_kernel(const double *__restrict__ in, double *__restrict__ out, long R, long X,
                                     int flip, long x_tiles)
{
    __shared__ double tile[32][33];
    const long x0 = ((long)blockIdx.x % x_tiles) * 32;
    const long r0 = ((long)blockIdx.x / x_tiles) * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const long r = r0 + q, x = x0 + threadIdx.x;
        tile[q][threadIdx.x] = (r < R && x < X) ? in[r * X + x] : 0.;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const long x = x0 + q, r = r0 + threadIdx.x;
        if (x < X && r < R) out[x * R + (flip ? R - 1 - r : r)] = tile[threadIdx.x][q];
    }
}

static void launch_transpose_rec(const double *d_in, double *d_out, long R, long X, int flip)
{
    const long x_tiles = (X + 31) / 32, r_tiles = (R + 31) / 32;
    QGSB_REQUIRE(x_tiles * r_tiles < (1L << 31), "record buffer too large for one layout-change launch");
    dim3 block(32, 8);
    transpose_rec_kernel<<<(unsigned)(x_tiles * r_tiles), block, 0, ctx().stream>>>(d_in, d_out, R, X, flip, x_tiles);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

// per-variable sum and sum of squares over members (ensemble statistics), deterministic two-stage reduction:
// block (i, r, p) sums variable i of record r over its slice of the members, the second kernel adds the slices
// in a fixed order.  Padding members (>= N) are skipped: they are integrated too and do not stay zero.
__global__ void moments_partial_kernel(const double *__restrict__ rec, long N, int n, long ld, int parts,
                                       double *__restrict__ partial)
{
    const int i = blockIdx.x, r = blockIdx.y, p = blockIdx.z;
    const double *y = rec + (size_t)r * n * ld;
    const long per = (N + parts - 1) / parts;
    const long lo = (long)p * per, hi = min(N, lo + per);
    double s1 = 0., s2 = 0.;
    for (long mbr = lo + threadIdx.x; mbr < hi; mbr += blockDim.x) {
        const double v = y[tile_base(mbr, n) + (size_t)i * TILE];
        s1 += v;
        s2 = fma(v, v, s2);
    }
    __shared__ double sh1[256], sh2[256];
    sh1[threadIdx.x] = s1;
    sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh1[threadIdx.x] += sh1[threadIdx.x + o];
            sh2[threadIdx.x] += sh2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *out = partial + ((size_t)r * n + i) * 2 * parts;
        out[p] = sh1[0];
        out[parts + p] = sh2[0];
    }
}

__global__ void moments_final_kernel(const double *__restrict__ partial, long count, int parts, double *__restrict__ sum,
                                     double *__restrict__ sumsq)
{
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (record, variable) pair
    if (q >= count) return;
    const double *in = partial + (size_t)q * 2 * parts;
    double s1 = 0., s2 = 0.;
    for (int p = 0; p < parts; ++p) {
        s1 += in[p];
        s2 += in[parts + p];
    }
    sum[q] = s1;
    sumsq[q] = s2;
}

// sums of R records (R, n, ld tiled) -> device arrays sum, sumsq (R, n)
void launch_record_moments(const double *d_rec, long R, long N, int n, long ld, double *d_sum, double *d_sumsq)
{
    if (R <= 0) return;
    const int parts = (int)std::max<long>(1, std::min<long>(64, N / 4096));
    PoolBuf<double> partial((size_t)R * n * 2 * parts);
    for (long r0 = 0; r0 < R; r0 += 65535) {
        const long rc = std::min<long>(65535, R - r0);
        dim3 grid((unsigned)n, (unsigned)rc, (unsigned)parts);
        moments_partial_kernel<<<grid, 256, 0, ctx().stream>>>(d_rec + (size_t)r0 * n * ld, N, n, ld, parts,
                                                              partial.p + (size_t)r0 * n * 2 * parts);
    }
    const long count = R * n;
    moments_final_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx().stream>>>(partial.p, count, parts, d_sum, d_sumsq);
    count_launch(2);
    QGSB_CUDA(cudaGetLastError());
}

// per-variable sum and sum of squares over members (ensemble statistics)
__global__ void moments_kernel(const double *__restrict__ y, long N, int n, double *__restrict__ out)
{
    const int i = blockIdx.x;
    double s1 = 0., s2 = 0.;
    for (long mbr = threadIdx.x; mbr < N; mbr += blockDim.x) {
        const double v = y[tile_base(mbr, n) + (size_t)i * TILE];
        s1 += v;
        s2 += v * v;
    }
    __shared__ double sh1[256], sh2[256];
    sh1[threadIdx.x] = s1;
    sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh1[threadIdx.x] += sh1[threadIdx.x + o];
            sh2[threadIdx.x] += sh2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[i] = sh1[0];
        out[n + i] = sh2[0];
    }
}

// ------------------------------------------------------------------------------------------------
static size_t tg_small_doubles(const qgsb_tensor *t, int m, int s, bool dense)
{
    const int n = t->view.n;
    return (size_t)(n + 1) + 2 * (size_t)n + (size_t)s * n + (dense ? DP_TILE_DOUBLES : t->view.jac.npos) +
           2 * (size_t)m + 64;
}

// Large bases run the product J @ X as a dense GEMM on the FP64 tensor cores (dense_product): decided here, per launch.
// QGSB_TGLS_DENSE = 0 / 1 forces the sparse / the dense product (A/B measurements, tests).
static void choose_product(const qgsb_tensor *t, TgParams &P)
{
    const char *env = getenv("QGSB_TGLS_DENSE");
    const bool dense = env ? env[0] != '0' : (t->view.n >= 64 && P.m >= 8);
    P.jd_ld = dense ? (int)round_up(t->view.n, 8) : 0;
}

static void fill_common(TgParams &P, const Tableau &tab, long N, int m, int adjoint, double inverse)
{
    memset(&P, 0, sizeof(P));
    P.n_members = N;
    P.m = m;
    P.s = tab.s;
    P.adjoint = adjoint ? 1 : 0;
    P.inverse = inverse;
    for (int i = 0; i < tab.s; ++i) {
        P.b[i] = tab.b[i];
        for (int j = 0; j < tab.s; ++j) P.a[i * tab.s + j] = tab.a[(size_t)i * tab.s + j];
    }
}

// decides shared vs global placement of the (s + 2) n m matrix block; returns dynamic smem bytes
static size_t place_matrices(const qgsb_tensor *t, TgParams &P, DevBuf<double> &scratch, size_t extra_mats)
{
    const int n = t->view.n;
    choose_product(t, P);
    const size_t small = tg_small_doubles(t, P.m, P.s, P.jd_ld != 0) * 8;
    const size_t mats = ((size_t)(P.s + 2) * n * P.m + extra_mats + (size_t)P.jd_ld * P.jd_ld) * 8;
    QGSB_REQUIRE(small + 1024 <= ctx().smem_optin, "model too large for the tangent-linear kernel (%zu bytes of state)",
                 small);
    if (small + mats <= ctx().smem_optin) {
        P.scratch = nullptr;
        return small + mats;
    }
    P.scratch_per_member = mats / 8;
    scratch.alloc(P.scratch_per_member * (size_t)P.n_members);
    P.scratch = scratch.p;
    return small;
}

template <typename K>
static void set_smem_attr(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        QGSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// picks the Benettin kernel for a filled parameter block (device pointers)
void benettin_dispatch(const qgsb_tensor *t, const Tableau &tab, TgParams &P, DevBuf<double> &scratch)
{
    const int m = P.m;
    cudaStream_t st = ctx().stream;
    if (pack_tangent_supported(t, tab, m)) {
        launch_pack_tangent(t, P, true);
    } else {
        const size_t bytes = place_matrices(t, P, scratch, 0);
        auto go = [&](auto kernel) {
            set_smem_attr(kernel, bytes);
            kernel<<<(unsigned)P.n_members, TG_THREADS, bytes, st>>>(t->view, P);
            count_launch();
        };
        if (t->view.rank == 5) {
            if (P.jd_ld) go(lyap_kernel<5, true>);
            else go(lyap_kernel<5, false>);
        } else {
            if (P.jd_ld) go(lyap_kernel<3, true>);
            else go(lyap_kernel<3, false>);
        }
    }
    QGSB_CUDA(cudaGetLastError());
}

void benettin_fill_common(TgParams &P, const Tableau &tab, long N, int m, int adjoint, double inverse)
{
    fill_common(P, tab, N, m, adjoint, inverse);
}

void launch_transpose_records(const double *d_in, double *d_out, long R, long inner, int flip)
{
    launch_transpose_rec(d_in, d_out, R, inner, flip);
}

}  // namespace qgsb

using namespace qgsb;

extern "C" {

// One member batch of qgsb_rk_tgls_integrate on the calling thread's device: enqueue() uploads, launches the tangent
// kernel and the record transposes, collect() brings the records to the caller's arrays -- called after the NEXT batch
// has been enqueued, so that the downloads hide behind its integration (see BenettinBatch).
struct TglsBatch {
    long N = 0, R = 0;
    int n = 0, m = 0;
    double *traj = nullptr, *fmat = nullptr;
    PoolBuf<double> d_y, d_fm, d_dt, d_ry, d_rf, d_oy, d_of;
    DevBuf<double> scratch;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;

    ~TglsBatch()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (done) cudaEventDestroy(done);
    }

    void enqueue(const qgsb_tensor *t, long N_, const double *ic, int m_, const double *tg_ic, long n_steps,
                 const double *dt, int s, const double *a, const double *b, long write_steps, int time_direction,
                 int adjoint, double inverse_sign, long R_, double *traj_, double *fmat_)
    {
        Context &cx = ctx();
        cudaStream_t st = cx.stream;
        const Tableau tab = make_tableau(s, a, b);
        N = N_;
        R = R_;
        n = t->view.n;
        m = m_;
        traj = traj_;
        fmat = fmat_;
        const size_t nm = (size_t)n * m;
        QGSB_CUDA(cudaEventCreate(&ev0));
        QGSB_CUDA(cudaEventCreate(&ev1));
        QGSB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        d_y.alloc((size_t)N * n);
        d_fm.alloc((size_t)N * nm);
        d_dt.alloc(std::max<long>(n_steps, 1));
        d_ry.alloc((size_t)R * N * n);
        d_rf.alloc((size_t)R * N * nm);
        d_oy.alloc((size_t)R * N * n);
        d_of.alloc((size_t)R * N * nm);
        d_y.upload(ic, (size_t)N * n, st);
        d_fm.upload(tg_ic, (size_t)N * nm, st);
        if (n_steps) d_dt.upload(dt, n_steps, st);
        TgParams P;
        fill_common(P, tab, N, m, adjoint, inverse_sign);
        P.n_steps = n_steps;
        P.dt = d_dt.p;
        P.write_steps = write_steps;
        P.n_records = R;
        P.y = d_y.p;
        P.fm = d_fm.p;
        P.rec_y = d_ry.p;
        P.rec_fm = d_rf.p;
        QGSB_CUDA(cudaEventRecord(ev0, st));
        if (pack_tangent_supported(t, tab, m)) {
            launch_pack_tangent(t, P, false);
        } else {
            const size_t bytes = place_matrices(t, P, scratch, 0);
            auto go = [&](auto kernel) {
                set_smem_attr(kernel, bytes);
                kernel<<<(unsigned)N, TG_THREADS, bytes, st>>>(t->view, P);
                count_launch();
            };
            if (t->view.rank == 5) {
                if (P.jd_ld) go(tgls_kernel<5, true>);
                else go(tgls_kernel<5, false>);
            } else {
                if (P.jd_ld) go(tgls_kernel<3, true>);
                else go(tgls_kernel<3, false>);
            }
        }
        QGSB_CUDA(cudaGetLastError());
        QGSB_CUDA(cudaEventRecord(ev1, st));
        launch_transpose_rec(d_ry.p, d_oy.p, R, (long)N * n, time_direction == -1);
        launch_transpose_rec(d_rf.p, d_of.p, R, (long)N * (long)nm, time_direction == -1);
        QGSB_CUDA(cudaEventRecord(done, st));
    }

    double collect()
    {
        cudaStream_t so = ctx().copy_out;
        QGSB_CUDA(cudaStreamWaitEvent(so, done, 0));
        d_oy.download(traj, (size_t)R * N * n, so);
        d_of.download(fmat, (size_t)R * N * n * m, so);
        QGSB_CUDA(cudaStreamSynchronize(so));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        return ms;
    }
};

// records kept on the device by one tangent / Benettin launch: bounded so that a long write_steps = 1 run is cut
// into member batches instead of failing in cudaMalloc (the reference keeps such runs in host RAM)
static long tangent_member_batch(long N, size_t bytes_per_member)
{
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return N;
    size_t budget = free_b / 2;
    if (const char *env = getenv("QGSB_TANGENT_BUDGET_MB")) budget = (size_t)atol(env) << 20;   // tests
    const long fit = (long)std::max<size_t>(1, budget / std::max<size_t>(bytes_per_member, 1));
    return std::min(N, fit);
}

int qgsb_rk_tgls_integrate(const qgsb_tensor *t, long N, const double *ic, int m, const double *tg_ic, long n_steps,
                           const double *dt, int s, const double *a, const double *b, const double *c,
                           long write_steps, int time_direction, int adjoint, double inverse_sign, long R,
                           double *traj, double *fmat, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && ic && tg_ic && traj && fmat, "null argument");
    QGSB_REQUIRE(N >= 1 && m >= 1, "need at least one trajectory and one tangent vector");
    QGSB_REQUIRE(n_steps >= 0 && write_steps >= 0, "negative step count");
    QGSB_REQUIRE(t->jnnz_in > 0, "tensor handle has no Jacobian tensor");
    QGSB_REQUIRE(time_direction == 1 || time_direction == -1, "time_direction must be +1 or -1");
    ensure_init();
    const int n = t->view.n;
    const size_t nm = (size_t)n * m;
    {
        long L = n_steps + 1, r = write_steps == 0 ? 1 : (L + write_steps - 1) / write_steps;
        if (write_steps > 0 && (r - 1) * write_steps != L - 1) r += 1;
        QGSB_REQUIRE(R == r, "n_records %ld inconsistent with %ld steps / write_steps %ld", R, n_steps, write_steps);
    }
    const int parts = shard_count(N, 1024);
    std::vector<double> ms(parts, 0.);
    run_sharded(N, parts, [&](int g, long lo, long hi) {
        const qgsb_tensor *th = tensor_here(t);
        // member batches sized to the device memory: records (R, batch, n + n m) twice (kernel order + API order), two
        // batches alive at a time; a large shard is cut into at least four (whole waves of the packed kernel) so that
        // the downloads of one batch hide behind the integration of the next
        long batch = tangent_member_batch(hi - lo, 2 * (size_t)(2 * R + 1) * (n + nm) * sizeof(double));
        const long wave = (long)ctx().sm_count * std::max(1, 256 / m);
        if (hi - lo >= 4 * wave) batch = std::min(batch, ((hi - lo + 3) / 4 + wave - 1) / wave * wave);
        std::unique_ptr<TglsBatch> previous;
        for (long m0 = lo; m0 < hi; m0 += batch) {
            const long nb = std::min(batch, hi - m0);
            std::unique_ptr<TglsBatch> current(new TglsBatch());
            current->enqueue(th, nb, ic + (size_t)m0 * n, m, tg_ic + (size_t)m0 * nm, n_steps, dt, s, a, b, write_steps,
                             time_direction, adjoint, inverse_sign, R, traj + (size_t)m0 * n * R,
                             fmat + (size_t)m0 * nm * R);
            if (previous) ms[g] += previous->collect();
            previous = std::move(current);
        }
        if (previous) ms[g] += previous->collect();
        QGSB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
    if (device_ms) *device_ms = *std::max_element(ms.begin(), ms.end());
    QGSB_API_END
}

// ---- start bases drawn on the device ------------------------------------------------------------------------------
// The reference draws qr(random((n_dim, n_vec))) per member from numba's generator (lyapunov.py:592-593).  With
// q0 == NULL the same is done here without host work: a counter-based generator (splitmix64 of seed, GLOBAL member
// index and element index -- so a member's draw does not depend on how the ensemble is split over devices or batches)
// fills the matrices with uniform [0, 1) numbers and the Benettin kernel factorises them with its own Householder QR
// before the first step (TgParams::qr_at_start).
static uint64_t g_seed = 0x243F6A8885A308D3ULL;
static long g_member_offset = 0;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void random_basis_kernel(double *__restrict__ q, long n_members, long per_member, long member0, uint64_t seed)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_members * per_member) return;
    const long member = idx / per_member, e = idx - member * per_member;
    const uint64_t h = splitmix64(splitmix64(seed ^ splitmix64((uint64_t)(member0 + member))) + (uint64_t)e);
    q[idx] = (double)(h >> 11) * (1. / 9007199254740992.);     // 53 random bits -> [0, 1)
}

// One member batch of qgsb_lyap_benettin on the calling thread's device; member0 = index of its first member in the whole
// ensemble.  enqueue() uploads, launches the Benettin kernel and the record transposes; collect() brings the records to
// the caller's arrays.  The caller enqueues batch k + 1 BEFORE it collects batch k: the downloads (which block the
// host thread when the destination is an ordinary pageable numpy array) then run on the copy stream while the device
// integrates the next batch.
struct BenettinBatch {
    long N = 0, steps = 0, n_rec = 0, R = 0;
    int n = 0, m = 0;
    double *rec_traj = nullptr, *rec_exp = nullptr, *rec_vec = nullptr, *r_all = nullptr, *q_all = nullptr;
    std::vector<long> f_ptr, idx;
    std::vector<double> f_sub, fdt;
    PoolBuf<double> d_y, d_q, d_dtm, d_sub, d_ry, d_rv, d_re, d_oy, d_ov, d_oe;
    PoolBuf<long> d_ptr, d_idx;
    DevBuf<double> d_r0, d_rall, d_qall, d_stored, d_fdt, d_state, scratch;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;

    ~BenettinBatch()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (done) cudaEventDestroy(done);
    }

    void enqueue(const qgsb_tensor *t, long N_, long member0, const double *ic, int forward, int n_vec, const double *q0,
                 const double *r0, long n_pre, long n_rec_, const double *dt_macro, const long *sub_ptr,
                 const double *sub_dt, int s, const double *a, const double *b, long write_steps, int adjoint,
                 double inverse_sign, long R_, double *rec_traj_, double *rec_exp_, double *rec_vec_, double *r_all_,
                 double *q_all_)
    {
        Context &cx = ctx();
        cudaStream_t st = cx.stream;
        const Tableau tab = make_tableau(s, a, b);
        N = N_;
        n = t->view.n;
        m = n_vec;
        n_rec = n_rec_;
        R = R_;
        rec_traj = rec_traj_;
        rec_exp = rec_exp_;
        rec_vec = rec_vec_;
        r_all = r_all_;
        q_all = q_all_;
        const size_t nm = (size_t)n * m;
        steps = n_pre + n_rec;
        QGSB_CUDA(cudaEventCreate(&ev0));
        QGSB_CUDA(cudaEventCreate(&ev1));
        QGSB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        // Micro steps of length exactly 0 are dropped: the reference's concatenate(arange(tt, tt + dt, mdt), [tt + dt])
        // (lyapunov.py:598) often ends in two equal times when mdt divides dt (375 of the 1000 steps of
        // arange(0, 100, 0.1)), and a Runge-Kutta step of length 0 leaves the state and the tangent matrix unchanged to
        // the last bit (x + 0 * k = x), while costing a full step -- and hiding that the step is a single micro step of
        // the macro length.
        f_ptr.assign(steps + 1, 0);
        f_sub.reserve(sub_ptr[steps]);
        for (long q = 0; q < steps; ++q) {
            for (long e = sub_ptr[q]; e < sub_ptr[q + 1]; ++e)
                if (sub_dt[e] != 0.) f_sub.push_back(sub_dt[e]);
            f_ptr[q + 1] = (long)f_sub.size();
        }
        const long n_sub = f_ptr[steps];
        d_y.alloc((size_t)N * n);
        d_q.alloc((size_t)N * nm);
        d_dtm.alloc(std::max<long>(steps, 1));
        d_sub.alloc(std::max<long>(n_sub, 1));
        d_ptr.alloc(steps + 1);
        d_idx.alloc(std::max<long>(steps, 1));
        // rec_vec == NULL: the vectors are not recorded (spectrum-only runs skip 8 n m bytes per member and record)
        d_ry.alloc((size_t)R * N * n);
        d_rv.alloc(rec_vec ? (size_t)R * N * nm : 0);
        d_re.alloc((size_t)R * N * m);
        d_oy.alloc((size_t)R * N * n);
        d_ov.alloc(rec_vec ? (size_t)R * N * nm : 0);
        d_oe.alloc((size_t)R * N * m);
        d_y.upload(ic, (size_t)N * n, st);
        if (q0) {
            d_q.upload(q0, (size_t)N * nm, st);
        } else {
            const long total = N * (long)nm;
            random_basis_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_q.p, N, (long)nm,
                                                                                 g_member_offset + member0, g_seed);
            count_launch();
            QGSB_CUDA(cudaGetLastError());
        }
        if (steps) d_dtm.upload(dt_macro, steps, st);
        if (n_sub) d_sub.upload(f_sub.data(), n_sub, st);
        QGSB_CUDA(cudaMemcpyAsync(d_ptr.p, f_ptr.data(), sizeof(long) * (steps + 1), cudaMemcpyHostToDevice, st));
        TgParams P;
        fill_common(P, tab, N, m, adjoint, inverse_sign);
        P.forward = forward;
        P.n_pre = n_pre;
        P.n_rec = n_rec;
        P.dt_macro = d_dtm.p;
        P.sub_ptr = d_ptr.p;
        P.sub_dt = d_sub.p;
        P.write_steps = write_steps;
        P.n_records = R;
        P.y = d_y.p;
        P.fm = d_q.p;
        P.rec_y = d_ry.p;
        P.rec_fm = rec_vec ? d_rv.p : nullptr;
        P.rec_exp = d_re.p;
        P.qr_at_start = q0 ? 0 : 1;
        if (r0 && q0) {
            d_r0.alloc((size_t)N * m * m);
            d_r0.upload(r0, (size_t)N * m * m, st);
            P.r0 = d_r0.p;
        }
        if (r_all) {
            d_rall.alloc((size_t)N * std::max<long>(steps, 1) * m * m);
            P.r_all = d_rall.p;
        }
        if (q_all) {
            d_qall.alloc((size_t)N * (n_rec + 1) * nm);
            P.q_all = d_qall.p;
        }
        QGSB_CUDA(cudaEventRecord(ev0, st));
        if (forward == 1) {
            // lyapunov.py:474: store the whole write_steps=1 trajectory, then walk it backwards.  The
            // steps come in execution (backward) order with negative dt; the forward integration uses
            // them reversed and positive.
            const long ld = round_up(N, TILE);
            fdt.resize(steps);
            for (long q = 0; q < steps; ++q) fdt[q] = -dt_macro[steps - 1 - q];
            d_fdt.alloc(std::max<long>(steps, 1));
            d_state.alloc((size_t)n * ld);
            if (steps) d_fdt.upload(fdt.data(), steps, st);
            d_stored.alloc((size_t)(steps + 1) * n * ld);
            launch_aos_to_soa(d_y.p, d_state.p, N, n, ld);
            rk_advance(t, d_state.p, ld, N, steps, d_fdt.p, tab, 1, steps + 1, d_stored.p);
            idx.resize(std::max<long>(steps, 1));
            for (long q = 0; q < steps; ++q) idx[q] = steps - q;  // step q starts from point steps - q
            QGSB_CUDA(cudaMemcpyAsync(d_idx.p, idx.data(), sizeof(long) * std::max<long>(steps, 1),
                                      cudaMemcpyHostToDevice, st));
            P.stored = d_stored.p;
            P.stored_ld = ld;
            P.start_idx = d_idx.p;
            P.final_idx = steps > 0 ? idx[steps - 1] : 0;   // lyapunov.py:549: y[0] of the last pass
        }
        benettin_dispatch(t, tab, P, scratch);
        QGSB_CUDA(cudaGetLastError());
        QGSB_CUDA(cudaEventRecord(ev1, st));
        launch_transpose_rec(d_ry.p, d_oy.p, R, (long)N * n, 0);
        if (rec_vec) launch_transpose_rec(d_rv.p, d_ov.p, R, (long)N * (long)nm, 0);
        launch_transpose_rec(d_re.p, d_oe.p, R, (long)N * m, 0);
        QGSB_CUDA(cudaEventRecord(done, st));
    }

    double collect()
    {
        Context &cx = ctx();
        cudaStream_t so = cx.copy_out;
        const size_t nm = (size_t)n * m;
        QGSB_CUDA(cudaStreamWaitEvent(so, done, 0));
        d_oy.download(rec_traj, (size_t)R * N * n, so);
        if (rec_vec) d_ov.download(rec_vec, (size_t)R * N * nm, so);
        d_oe.download(rec_exp, (size_t)R * N * m, so);
        if (r_all) d_rall.download(r_all, (size_t)N * steps * m * m, so);
        if (q_all) d_qall.download(q_all, (size_t)N * (n_rec + 1) * nm, so);
        QGSB_CUDA(cudaStreamSynchronize(so));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        return ms;
    }
};

int qgsb_set_seed(uint64_t seed, long member_offset)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(member_offset >= 0, "negative member offset");
    g_seed = seed;
    g_member_offset = member_offset;
    QGSB_API_END
}

int qgsb_lyap_benettin(const qgsb_tensor *t, long N, const double *ic, int forward, int n_vec, const double *q0,
                       const double *r0, long n_pre, long n_rec, const double *dt_macro, const long *sub_ptr,
                       const double *sub_dt, int s, const double *a, const double *b, const double *c,
                       long write_steps, int adjoint, double inverse_sign, long R, double *rec_traj,
                       double *rec_exp, double *rec_vec, double *r_all, double *q_all, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && ic && dt_macro && sub_ptr && sub_dt && rec_traj && rec_exp, "null argument");
    QGSB_REQUIRE(N >= 1, "need at least one trajectory");
    QGSB_REQUIRE(forward >= 0 && forward <= 2, "mode must be 0 (BLV), 1 (FLV) or 2 (BLV following the micro-steps)");
    QGSB_REQUIRE(n_vec >= 1 && n_vec <= t->view.n, "n_vec must be in 1..n_dim");
    QGSB_REQUIRE(n_pre >= 0 && n_rec >= 0 && write_steps >= 0, "negative step count");
    QGSB_REQUIRE(t->jnnz_in > 0, "tensor handle has no Jacobian tensor");
    ensure_init();
    const int n = t->view.n, m = n_vec;
    const size_t nm = (size_t)n * m;
    const long steps = n_pre + n_rec;
    {
        long L = n_rec + 1, r = write_steps == 0 ? 1 : (L + write_steps - 1) / write_steps;
        if (write_steps > 0 && (r - 1) * write_steps != L - 1) r += 1;
        QGSB_REQUIRE(R == r, "n_records %ld inconsistent with %ld recorded steps / write_steps %ld", R, n_rec,
                     write_steps);
    }
    // bytes one member keeps on the device during the launch: records in kernel and API order, plus the optional
    // stored trajectory (FLV), R factors and bases (Ginelli)
    const size_t per_member = sizeof(double) * ((size_t)2 * R * (n + m + (rec_vec ? nm : 0)) + n + nm +
                                                (forward == 1 ? (size_t)(steps + 1) * n : 0) +
                                                (r_all ? (size_t)std::max<long>(steps, 1) * m * m : 0) +
                                                (q_all ? (size_t)(n_rec + 1) * nm : 0));
    const int parts = shard_count(N, 1024);
    std::vector<double> ms(parts, 0.);
    run_sharded(N, parts, [&](int g, long lo, long hi) {
        const qgsb_tensor *th = tensor_here(t);
        // Member batches: as many as device memory demands (two batches are alive at a time), and for a large shard at
        // least four, cut at whole waves of the packed kernel, so that the record downloads of one batch hide behind
        // the integration of the next.
        long batch = tangent_member_batch(hi - lo, 2 * per_member);
        const long wave = (long)ctx().sm_count * std::max(1, 256 / m);
        if (hi - lo >= 4 * wave) batch = std::min(batch, ((hi - lo + 3) / 4 + wave - 1) / wave * wave);
        std::unique_ptr<BenettinBatch> previous;
        for (long m0 = lo; m0 < hi; m0 += batch) {
            const long nb = std::min(batch, hi - m0);
            std::unique_ptr<BenettinBatch> current(new BenettinBatch());
            current->enqueue(th, nb, m0, ic + (size_t)m0 * n, forward, n_vec, q0 ? q0 + (size_t)m0 * nm : nullptr,
                             r0 ? r0 + (size_t)m0 * m * m : nullptr, n_pre, n_rec, dt_macro, sub_ptr, sub_dt, s, a, b,
                             write_steps, adjoint, inverse_sign, R, rec_traj + (size_t)m0 * n * R,
                             rec_exp + (size_t)m0 * m * R, rec_vec ? rec_vec + (size_t)m0 * nm * R : nullptr,
                             r_all ? r_all + (size_t)m0 * steps * m * m : nullptr,
                             q_all ? q_all + (size_t)m0 * (n_rec + 1) * nm : nullptr);
            if (previous) ms[g] += previous->collect();
            previous = std::move(current);
        }
        if (previous) ms[g] += previous->collect();
        QGSB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
    if (device_ms) *device_ms = *std::max_element(ms.begin(), ms.end());
    QGSB_API_END
}

static void ensemble_moments_device(qgsb_ensemble *e, double *sum, double *sumsq)
{
    const int n = e->tensor->view.n;
    PoolBuf<double> d_out((size_t)2 * n);
    launch_record_moments(e->d_y.p, 1, e->N, n, e->ld, d_out.p, d_out.p + n);
    d_out.download(sum, n, ctx().stream);
    QGSB_CUDA(cudaMemcpyAsync(sumsq, d_out.p + n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx().stream));
    QGSB_CUDA(cudaStreamSynchronize(ctx().stream));
}

int qgsb_ensemble_moments(qgsb_ensemble *e, double *sum, double *sumsq)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(e && sum && sumsq, "null argument");
    ensure_init();
    if (e->parts.empty()) {
        ensemble_moments_device(e, sum, sumsq);
    } else {
        // per-part sums on the parts' devices, added in part order (deterministic)
        const int n = e->tensor->view.n, parts = (int)e->parts.size();
        std::vector<double> s1((size_t)n * parts), s2((size_t)n * parts);
        run_sharded(e->N, parts, [&](int g, long, long) {
            ensemble_moments_device(e->parts[g], s1.data() + (size_t)n * g, s2.data() + (size_t)n * g);
        });
        for (int i = 0; i < n; ++i) {
            double t1 = 0., t2 = 0.;
            for (int g = 0; g < parts; ++g) {
                t1 += s1[(size_t)n * g + i];
                t2 += s2[(size_t)n * g + i];
            }
            sum[i] = t1;
            sumsq[i] = t2;
        }
    }
    QGSB_API_END
}

}  // extern "C"
