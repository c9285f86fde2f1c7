// tgls_pack.cuh -- "packed" tangent-linear / Benettin kernels for small bases (ndim <= ~40).
//
// Same arithmetic as tgls.cu (integrate.py:555-614, lyapunov.py:471-632), laid out for the FP64 pipe:
//
//  * a thread owns ONE COLUMN of the n x m tangent matrix of one member in registers (ndim is a
//    template parameter, every register index is a literal);
//  * a block packs G = floor(256 / m) members, so all lanes of (almost) every warp carry a column --
//    a member is not tied to a warp or a block, only to m consecutive threads (consecutive, so that the
//    broadcast reads of a member's J and reflectors hit one or two addresses per warp; interleaving the
//    members across lanes was measured 20 % slower);
//  * per Runge-Kutta stage the m threads of a member evaluate the tendencies and the values of the
//    structurally non-zero Jacobian positions of THEIR member into shared memory, then every thread
//    computes km[:, c] = +-J @ kms[:, c] reading J as broadcast LDS.128 (two positions per load).
//    The product is a policy: `DenseProduct` walks a dense n x n matrix (any tensor), a generated
//    module (qgs_b200/codegen.py) supplies the same product as straight-line code over the literal
//    position list of one tensor (MAOOAM-36: 490 positions instead of 1296);
//  * the Benettin re-orthonormalisation is a Householder QR with LAPACK's dgeqr2 / dorg2r sign
//    conventions (so Q, R match np.linalg.qr) WITHOUT a serial phase: the owner of column j publishes
//    its raw column, every later column thread forms the dot product with it AND its norm (so beta,
//    tau and the scaling are computed redundantly by everybody, no second barrier), and the explicit
//    Q is accumulated from the published reflectors with no barrier at all.
//
// Only "chain" tableaux (a_ij != 0 only for j = i-1: Euler, midpoint, Heun, classic RK4) take this
// path; the generic kernels of tgls.cu serve everything else.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "tgls_shared.cuh"

namespace qgsb {
namespace pack {

#ifndef QGSB_PACK_THREADS
#define QGSB_PACK_THREADS 256
#endif
// 1: a lean build for modules compiled at run time (codegen.build_plugin): the rolled Householder kernels only (bitwise
// the unrolled ones, a few per cent slower on full bases) and the Cholesky QR for the full basis only (partial bases
// keep Householder).  The full set takes nvcc two minutes per module, the lean one half a minute.
#ifndef QGSB_PACK_LEAN
#define QGSB_PACK_LEAN 0
#endif
#ifndef QGSB_PACK_BLOCKS
#define QGSB_PACK_BLOCKS 1
#endif
constexpr int MAX_THREADS = QGSB_PACK_THREADS;
constexpr int MIN_BLOCKS = QGSB_PACK_BLOCKS;      // resident blocks per SM the register allocation is sized for

// launch geometry decided on the host
struct Geometry {
    int G = 0;            // members per block
    int threads = 0;      // G * m rounded up to a warp
    int stride = 0;       // doubles of shared memory per member
    int jv = 0;           // doubles reserved for the Jacobian values of one member
    size_t smem = 0;      // dynamic shared memory per block, bytes
};

__host__ __device__ inline int even(int x) { return (x + 1) & ~1; }

// per-member carve-up of shared memory (offsets in doubles, all even => 16-byte aligned)
template <int N>
struct Carve {
    int jv, mp, nm;
    __host__ __device__ Carve(int jv_, int m) : jv(even(jv_)), mp(even(m)), nm(even(N * m)) {}
    __host__ __device__ int o_jv() const { return 0; }
    __host__ __device__ int o_xs() const { return jv; }                 // N + 2 (xs[0] = 1)
    __host__ __device__ int o_xs2() const { return o_xs() + even(N + 2); }       // second stage-state buffer
    __host__ __device__ int o_y() const { return o_xs2() + even(N + 2); }
    __host__ __device__ int o_Y() const { return o_y() + even(N); }
    __host__ __device__ int o_kst() const { return o_Y() + even(N); }
    __host__ __device__ int o_yacc() const { return o_kst() + even(N); }
    __host__ __device__ int o_rdiag() const { return o_yacc() + even(N); }
    __host__ __device__ int o_tau() const { return o_rdiag() + mp; }
    __host__ __device__ int o_scal() const { return o_tau() + mp; }
    __host__ __device__ int o_fm() const { return o_scal() + mp; }
    __host__ __device__ int o_facc() const { return o_fm() + nm; }
    __host__ __device__ int total() const
    {
        int t = o_facc() + nm;
        // members of one warp read the same literal offset of their own J: keep their bases in
        // different bank groups (stride = 2 mod 16 doubles)
        while (t % 16 != 2) t += 2;
        return t;
    }
};

template <int N>
struct Mem {
    double *jv, *xs, *xs2, *y, *Y, *kst, *yacc, *rdiag, *tau, *scal, *fm, *facc;
    int m;
};

template <int N>
__device__ __forceinline__ Mem<N> carve(double *base, int jv, int m)
{
    const Carve<N> c(jv, m);
    Mem<N> S;
    S.jv = base + c.o_jv();
    S.xs = base + c.o_xs();
    S.xs2 = base + c.o_xs2();
    S.y = base + c.o_y();
    S.Y = base + c.o_Y();
    S.kst = base + c.o_kst();
    S.yacc = base + c.o_yacc();
    S.rdiag = base + c.o_rdiag();
    S.tau = base + c.o_tau();
    S.scal = base + c.o_scal();
    S.fm = base + c.o_fm();
    S.facc = base + c.o_facc();
    S.m = m;
    return S;
}

// ---- product policies ------------------------------------------------------------------------------------------
// A policy provides
//   JV              doubles of shared memory per member for the Jacobian values
//   kUsesJacobianValues   false: apply() works from the stage state xs directly (generated bilinear form)
//   slot(i, j)      where the value of position (i, j) (1-based) goes inside that area
//   apply(jv, xs, col, km)   km = (J or J^T) @ col
template <int N, bool ADJ>
struct DenseProduct {
    static constexpr int JV = N * N;
    static constexpr bool kUsesJacobianValues = true;
    static constexpr bool kZeroFill = true;   // structural zeros must read as 0
    __device__ static __forceinline__ int slot(int i, int j) { return (i - 1) * N + (j - 1); }
    __device__ static __forceinline__ void apply(const double *jv, const double *, const double (&col)[N],
                                                 double (&km)[N])
    {
#pragma unroll
        for (int i = 0; i < N; ++i) km[i] = 0.;
        if (!ADJ) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double2 *row = reinterpret_cast<const double2 *>(jv + i * N);
#pragma unroll
                for (int j = 0; j < N / 2; ++j) {
                    const double2 v = row[j];
                    km[i] = fma(v.x, col[2 * j], km[i]);
                    km[i] = fma(v.y, col[2 * j + 1], km[i]);
                }
                if (N & 1) km[i] = fma(jv[i * N + N - 1], col[N - 1], km[i]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 *row = reinterpret_cast<const double2 *>(jv + j * N);
#pragma unroll
                for (int i = 0; i < N / 2; ++i) {
                    const double2 v = row[i];
                    km[2 * i] = fma(v.x, col[j], km[2 * i]);
                    km[2 * i + 1] = fma(v.y, col[j], km[2 * i + 1]);
                }
                if (N & 1) km[N - 1] = fma(jv[j * N + N - 1], col[j], km[N - 1]);
            }
        }
    }
};

// ---- tendencies and Jacobian values of one member ------------------------------------------------------------------
// The ELL tables live in shared memory (ShTab: plain shared pointers, so every access below is an LDS); their index
// fields are BYTE offsets into the augmented stage state, and entry 0 of a Jacobian position carries the byte offset
// of its slot in the member's Jacobian area.
struct ShTab {
    const PEnt *f = nullptr;   // (EF, N)
    const PEnt *j = nullptr;   // (EJ, npos)
    int EF = 0, EJ = 0, npos = 0;
};

__device__ __forceinline__ double at(const double *base, unsigned off)
{
    return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + off);
}

// row r of f at the stage state xs (sparse_mul.py:76-81 / :153-158 with the row's entries in COO order)
template <int N>
__device__ __forceinline__ double f_row_tab(const TensorView &T, const ShTab &tab, int r, const double *xs)
{
    if (tab.f == nullptr) return f_row_rt(T, r + 1, xs);
    const PEnt *e = tab.f + r;
    double acc = 0.;
    if (T.rank == 5) {
#pragma unroll 2
        for (int q = 0; q < tab.EF; ++q) {
            const PEnt en = e[q * N];
            acc += at(xs, en.a) * at(xs, en.b) * at(xs, en.c) * at(xs, en.d) * en.v;
        }
    } else {
        // the row's own length (spare field of its entry 0): the lanes of a warp hold different rows, and the loads
        // of a lane that is done are not issued -- shared-memory traffic follows the 351 real entries of MAOOAM-36
        // instead of 36 rows x 15
        const int len = e[0].d;
#pragma unroll 5
        for (int q = 0; q < len; ++q) {
            const PEnt en = e[q * N];
            acc += at(xs, en.a) * at(xs, en.b) * en.v;
        }
    }
    return acc;
}

// values of the Jacobian positions q = c, c + m, ... of this member (sparse_mul.py:40-44 / :113-117)
template <int N, class Prod>
__device__ __forceinline__ void jac_build(const TensorView &T, const ShTab &tab, const double *xs, double *jv, int c,
                                          int m)
{
    const JacView &J = T.jac;
    if (tab.j == nullptr) {
        for (int p = c; p < J.npos; p += m) jv[Prod::slot(J.pos_i[p], J.pos_j[p])] = jac_pos_rt(J, T.rank, p, xs);
        return;
    }
    const int npos = tab.npos;
    char *jvb = reinterpret_cast<char *>(jv);
    if (T.rank == 5) {
        for (int q = c; q < npos; q += m) {
            const unsigned slot = tab.j[q].d;
            double acc = 0.;
            for (int e = 0; e < tab.EJ; ++e) {
                const PEnt en = tab.j[e * npos + q];
                acc += at(xs, en.a) * at(xs, en.b) * at(xs, en.c) * en.v;
            }
            *reinterpret_cast<double *>(jvb + slot) = acc;
        }
    } else if (tab.EJ == 2) {
#pragma unroll 4
        for (int q = c; q < npos; q += m) {
            const PEnt e0 = tab.j[q], e1 = tab.j[npos + q];
            double acc = at(xs, e0.a) * e0.v;
            acc += at(xs, e1.a) * e1.v;
            *reinterpret_cast<double *>(jvb + e0.d) = acc;
        }
    } else {
        for (int q = c; q < npos; q += m) {
            const unsigned slot = tab.j[q].d;
            double acc = 0.;
            for (int e = 0; e < tab.EJ; ++e) {
                const PEnt en = tab.j[e * npos + q];
                acc += at(xs, en.a) * en.v;
            }
            *reinterpret_cast<double *>(jvb + slot) = acc;
        }
    }
}

// copies the tables into shared memory (after the members' areas)
__device__ __forceinline__ ShTab stage_tables(const PackTables &tab, double *dst, int n)
{
    ShTab out;
    PEnt *f = reinterpret_cast<PEnt *>(dst);
    const int nf = tab.f_ent ? tab.EF * n : 0, nj = tab.j_ent ? tab.EJ * tab.npos : 0;
    PEnt *j = f + nf;
    for (int q = threadIdx.x; q < nf; q += blockDim.x) f[q] = tab.f_ent[q];
    for (int q = threadIdx.x; q < nj; q += blockDim.x) j[q] = tab.j_ent[q];
    if (tab.f_ent) {
        out.f = f;
        out.EF = tab.EF;
    }
    if (tab.j_ent) {
        out.j = j;
        out.EJ = tab.EJ;
        out.npos = tab.npos;
    }
    return out;
}

// ---- one step of the coupled system (chain tableau) -----------------------------------------------------------------
// col[] holds fm[:, c] on entry and on exit; S.y advances by dt.  No barrier at the end: the caller
// synchronises before anybody reads another thread's data.
// FAREG: the weighted sum of the stage derivatives of the column stays in registers (fa) instead of in the column's
// slice of S.facc.  In shared memory the sum costs a load and a store per element and stage, a third of the stage
// epilogue's traffic; in registers it is 2 N more of them next to the column and the product.  [B200, 8192 members]
// tangent-linear kernel: MAOOAM-36 5.42 -> 6.11e7 member-steps/s, RP-20 1.24 -> 1.38e8; Benettin kernel: RP-20 6.36 ->
// 6.85e7, but MAOOAM-36 2.49 -> 2.40e7 (the Benettin loop keeps more state live across the step and spills 688 bytes
// at 36 variables): the plain integration always takes it, the Benettin loop only for small bases (N <= 24).
template <int N, class Prod, bool FAREG>
__device__ __forceinline__ void tangent_step(const TensorView &T, const ShTab &tab, const TgParams &P,
                                             const Mem<N> &S, double dt, double (&col)[N], int c, bool live,
                                             const Mem<N> &Sr, int cr, bool liver)
{
    // (Sr, cr, liver): the member and rows r = cr, cr + m, ... whose tendencies this thread evaluates.  They need not
    // be the thread's own member: dealt member-fastest, the lanes of a warp read the SAME table entries and gather
    // from the staggered state areas of different members (fewer bank conflicts than 32 different rows of one state).
    // Every access to y, yacc, kst and the row part of xs goes through this mapping, so no barrier is needed between
    // them.
    const int s = P.s, m = S.m;
    double km[N], fa[FAREG ? N : 1];
    for (int st = 0; st < s; ++st) {
        const double wa_in = st > 0 ? dt * P.a[st * s + st - 1] : 0.;       // (dt a[st]) @ k   integrate.py:216
        const double wb = dt * P.b[st];
        const double wa_out = st + 1 < s ? dt * P.a[(st + 1) * s + st] : 0.;
        // The stage state alternates between two buffers: a thread that is already forming the state of the next
        // stage cannot disturb a thread that still reads this one.  Products that work from the stage state itself
        // (generated bilinear form) then need ONE barrier per stage; those that read Jacobian values built by the
        // whole member need a second one.
        double *xs = (st & 1) ? S.xs2 : S.xs;
        double *xr = (st & 1) ? Sr.xs2 : Sr.xs;
        if (liver)
            for (int r = cr; r < N; r += m) xr[r + 1] = st > 0 ? Sr.y[r] + wa_in * Sr.kst[r] : Sr.y[r];
        __syncthreads();
        if (liver)
            for (int r = cr; r < N; r += m) {
                const double k = f_row_tab<N>(T, tab, r, xr);
                Sr.kst[r] = k;
                Sr.yacc[r] = st == 0 ? wb * k : Sr.yacc[r] + wb * k;
            }
        if (live && Prod::kUsesJacobianValues) jac_build<N, Prod>(T, tab, xs, S.jv, c, m);
        if (Prod::kUsesJacobianValues) __syncthreads();
        if (live) {
            // km = inverse * (J or J^T) @ col        integrate.py:601-603, boundary == 0
            Prod::apply(S.jv, xs, col, km);
            double *fmc = S.fm + c;
            double *fc = S.facc + c;
            if (st + 1 < s) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double k = P.inverse * km[i];
                    if (FAREG) {
                        fa[i] = st == 0 ? wb * k : fa[i] + wb * k;               // sum dt b_j km_j  :605-607
                    } else {
                        const double f = st == 0 ? wb * k : fc[i * m] + wb * k;
                        fc[i * m] = f;
                    }
                    col[i] = fmc[i * m] + wa_out * k;                            // km_s of the next stage :598-600
                }
            } else {
                // last stage: the sum is complete, the new column is fm + sum -- the same operations in the same
                // order as storing the sum and adding it in a second sweep, without that sweep's loads and stores
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double k = P.inverse * km[i];
                    const double f = st == 0 ? wb * k : (FAREG ? fa[i] : fc[i * m]) + wb * k;
                    col[i] = fmc[i * m] + f;
                    fmc[i * m] = col[i];
                }
            }
        }
    }
    if (liver)
        for (int r = cr; r < N; r += m) Sr.y[r] += Sr.yacc[r];
    // an odd number of stages ends on the buffer the next step starts with
    if (s & 1) __syncthreads();
}

// one nonlinear step of the macro ("stored") trajectory: S.Y <- RK(S.Y, dt)
template <int N>
__device__ __forceinline__ void nl_step(const TensorView &T, const ShTab &tab, const TgParams &P,
                                        const Mem<N> &S, double dt, int c, bool live)
{
    const int s = P.s, m = S.m;
    for (int st = 0; st < s; ++st) {
        const double wa_in = st > 0 ? dt * P.a[st * s + st - 1] : 0.;
        const double wb = dt * P.b[st];
        if (live)
            for (int r = c; r < N; r += m) S.xs[r + 1] = st > 0 ? S.Y[r] + wa_in * S.kst[r] : S.Y[r];
        __syncthreads();
        if (live)
            for (int r = c; r < N; r += m) {
                const double k = f_row_tab<N>(T, tab, r, S.xs);
                S.kst[r] = k;
                S.yacc[r] = st == 0 ? wb * k : S.yacc[r] + wb * k;
            }
        __syncthreads();
    }
    if (live)
        for (int r = c; r < N; r += m) S.Y[r] += S.yacc[r];
    __syncthreads();
}

// 1 / x to within 1 ulp for normal x of moderate magnitude: hardware seed + two Newton steps, without the
// branchy slow path of the IEEE division (the reflector scalars tolerate the last bit)
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.);
    r = fma(r, e, r);
    e = fma(-x, r, 1.);
    r = fma(r, e, r);
    e = fma(-x, r, 1.);
    return fma(r, e, r);
}

// Householder QR of the n x m matrix whose column c lives in col[] of thread c of the member.  On exit col[]
// holds Q[:, c] (also stored to S.fm), S.rdiag the diagonal of R, Rout (m x m row-major, may be null) all of R.
// The reflectors are kept column-major in the facc area: V[j * N + i] = raw x_i of step j (i >= j); their LAPACK
// scaling 1 / (alpha - beta) is kept apart in S.scal so that nobody rewrites a published column.
template <int N>
__device__ __forceinline__ void qr(const Mem<N> &S, int c, bool live, double (&col)[N], double *Rout)
{
    static_assert(N % 2 == 0, "rows are processed in aligned pairs");
    const int m = S.m;
    double *V = S.facc;
    __syncthreads();                       // every thread is done with its private facc column
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (j < m) {                       // uniform
            double *x = V + j * N;
            if (live && c == j) {
                if (j & 1) x[j] = col[j];
#pragma unroll
                for (int i = (j + 1) & ~1; i < N; i += 2)      // 16-byte stores; for even j the pair starts at row j
                    *reinterpret_cast<double2 *>(x + i) = make_double2(col[i], col[i + 1]);
            }
            __syncthreads();
            if (live && c >= j) {
                // dlarfg + dlarf: d = x . col_c and |x|^2 over the rows below the diagonal (4 + 4 independent chains)
                double d0 = 0., d1 = 0., d2 = 0., d3 = 0., n0 = 0., n1 = 0., n2 = 0., n3 = 0.;
#pragma unroll
                for (int i = (j + 1) & ~1; i < N; i += 2) {       // rows in aligned pairs: one LDS.128 each
                    const double2 xi = *reinterpret_cast<const double2 *>(x + i);
                    if ((i >> 1) & 1) {
                        if (i > j) {
                            d0 = fma(xi.x, col[i], d0);
                            n0 = fma(xi.x, xi.x, n0);
                        }
                        d1 = fma(xi.y, col[i + 1], d1);
                        n1 = fma(xi.y, xi.y, n1);
                    } else {
                        if (i > j) {
                            d2 = fma(xi.x, col[i], d2);
                            n2 = fma(xi.x, xi.x, n2);
                        }
                        d3 = fma(xi.y, col[i + 1], d3);
                        n3 = fma(xi.y, xi.y, n3);
                    }
                }
                const double alpha = x[j];
                const double nrm2 = (n0 + n1) + (n2 + n3), dot = (d0 + d1) + (d2 + d3);
                double beta = alpha, tau = 0., scal = 0.;
                if (nrm2 != 0.) {
                    // dlapy2(alpha, xnorm) = sqrt(alpha^2 + |x|^2): the columns are O(1) here (an orthonormal
                    // basis propagated over one step), so the unscaled form neither overflows nor underflows
                    beta = -copysign(sqrt(fma(alpha, alpha, nrm2)), alpha);
                    tau = (beta - alpha) * fast_rcp(beta);
                    scal = fast_rcp(alpha - beta);
                }
                if (c == j) {
                    col[j] = beta;
                    S.rdiag[j] = beta;
                    S.tau[j] = tau;
                    S.scal[j] = scal;
                } else {
                    const double w = tau * fma(dot, scal, col[j]);
                    col[j] -= w;
                    const double ws = -(w * scal);
#pragma unroll
                    for (int i = (j + 1) & ~1; i < N; i += 2) {
                        const double2 xi = *reinterpret_cast<const double2 *>(x + i);
                        if (i > j) col[i] = fma(ws, xi.x, col[i]);
                        col[i + 1] = fma(ws, xi.y, col[i + 1]);
                    }
                }
                if (Rout != nullptr) Rout[j * m + c] = col[j];
            }
        }
    }
    if (Rout != nullptr && live)
        for (int i = c + 1; i < m; ++i) Rout[i * m + c] = 0.;   // strictly lower part of column c
    __syncthreads();                       // the last reflector, tau and scal are published
    // dorg2r: Q[:, c] = H_0 ... H_{m-1} e_c; H_j e_c = e_c for j > c
#pragma unroll
    for (int i = 0; i < N; ++i) col[i] = i == c ? 1. : 0.;
#pragma unroll
    for (int jj = 0; jj < N; ++jj) {
        const int j = N - 1 - jj;
        if (j < m && live && j <= c) {
            const double *x = V + j * N;
            double d0 = 0., d1 = 0., d2 = 0., d3 = 0.;
#pragma unroll
            for (int i = (j + 1) & ~1; i < N; i += 2) {
                const double2 xi = *reinterpret_cast<const double2 *>(x + i);
                if ((i >> 1) & 1) {
                    if (i > j) d0 = fma(xi.x, col[i], d0);
                    d1 = fma(xi.y, col[i + 1], d1);
                } else {
                    if (i > j) d2 = fma(xi.x, col[i], d2);
                    d3 = fma(xi.y, col[i + 1], d3);
                }
            }
            const double scal = S.scal[j];
            const double w = S.tau[j] * fma((d0 + d1) + (d2 + d3), scal, col[j]);
            col[j] -= w;
            const double ws = -(w * scal);
#pragma unroll
            for (int i = (j + 1) & ~1; i < N; i += 2) {
                const double2 xi = *reinterpret_cast<const double2 *>(x + i);
                if (i > j) col[i] = fma(ws, xi.x, col[i]);
                col[i + 1] = fma(ws, xi.y, col[i + 1]);
            }
        }
    }
    if (live) {
        double *fmc = S.fm + c;
#pragma unroll
        for (int i = 0; i < N; ++i) fmc[i * m] = col[i];
    }
}

// ---- the same factorisation with the reflector loop ROLLED ------------------------------------------------------------
// The fully unrolled qr<N> above is ~11.6k instructions (186 KB) that every warp streams from L2 once per step -- ncu
// shows the Benettin kernel waiting for instruction fetch there (stall "no instruction" 1.1 per issue).  Here the
// reflectors are taken in three groups; inside a group j is a run-time index, rows start at the group's first row and
// the rows <= j are masked to zero, so one loop body per group (a few hundred instructions) serves all its reflectors.
// More FMAs (masked rows), far fewer instruction bytes.  Same arithmetic on the unmasked rows, same results.
template <int N, int JB, int JE>
__device__ __forceinline__ void qr_factor_group(const Mem<N> &S, int c, bool live, double (&col)[N], double *Rout)
{
    constexpr int RB = JB & ~1;
    const int m = S.m;
    double *V = S.facc;
    const int jend = JE < m ? JE : m;
#pragma unroll 1
    for (int j = JB; j < jend; ++j) {
        double *x = V + j * N;
        if (live && c == j) {
#pragma unroll
            for (int i = RB; i < N; i += 2) *reinterpret_cast<double2 *>(x + i) = make_double2(col[i], col[i + 1]);
        }
        __syncthreads();
        if (live && c >= j) {
            double d0 = 0., d1 = 0., d2 = 0., d3 = 0., n0 = 0., n1 = 0., n2 = 0., n3 = 0.;
#pragma unroll
            for (int i = RB; i < N; i += 2) {
                double2 xi = *reinterpret_cast<const double2 *>(x + i);
                if (i < JE) {                                  // rows that can be <= j in this group
                    if (i <= j) xi.x = 0.;
                    if (i + 1 <= j) xi.y = 0.;
                }
                if ((i >> 1) & 1) {
                    d0 = fma(xi.x, col[i], d0);
                    n0 = fma(xi.x, xi.x, n0);
                    d1 = fma(xi.y, col[i + 1], d1);
                    n1 = fma(xi.y, xi.y, n1);
                } else {
                    d2 = fma(xi.x, col[i], d2);
                    n2 = fma(xi.x, xi.x, n2);
                    d3 = fma(xi.y, col[i + 1], d3);
                    n3 = fma(xi.y, xi.y, n3);
                }
            }
            const double alpha = x[j];
            const double nrm2 = (n0 + n1) + (n2 + n3), dot = (d0 + d1) + (d2 + d3);
            double beta = alpha, tau = 0., scal = 0.;
            if (nrm2 != 0.) {
                beta = -copysign(sqrt(fma(alpha, alpha, nrm2)), alpha);
                tau = (beta - alpha) * fast_rcp(beta);
                scal = fast_rcp(alpha - beta);
            }
            double cj = 0.;
#pragma unroll
            for (int i = JB; i < JE; ++i)
                if (i == j) cj = col[i];
            if (c == j) {
                cj = beta;
                S.rdiag[j] = beta;
                S.tau[j] = tau;
                S.scal[j] = scal;
            } else {
                const double w = tau * fma(dot, scal, cj);
                cj -= w;
                const double ws = -(w * scal);
#pragma unroll
                for (int i = RB; i < N; i += 2) {
                    double2 xi = *reinterpret_cast<const double2 *>(x + i);
                    if (i < JE) {
                        if (i <= j) xi.x = 0.;
                        if (i + 1 <= j) xi.y = 0.;
                    }
                    col[i] = fma(ws, xi.x, col[i]);
                    col[i + 1] = fma(ws, xi.y, col[i + 1]);
                }
            }
#pragma unroll
            for (int i = JB; i < JE; ++i)
                if (i == j) col[i] = cj;
            if (Rout != nullptr) Rout[j * m + c] = cj;
        }
    }
}

template <int N, int JB, int JE>
__device__ __forceinline__ void qr_formq_group(const Mem<N> &S, int c, bool live, double (&col)[N])
{
    constexpr int RB = JB & ~1;
    const int m = S.m;
    const double *V = S.facc;
    const int jend = JE < m ? JE : m;
#pragma unroll 1
    for (int j = jend - 1; j >= JB; --j) {
        if (live && j <= c) {
            const double *x = V + j * N;
            double d0 = 0., d1 = 0., d2 = 0., d3 = 0.;
#pragma unroll
            for (int i = RB; i < N; i += 2) {
                double2 xi = *reinterpret_cast<const double2 *>(x + i);
                if (i < JE) {
                    if (i <= j) xi.x = 0.;
                    if (i + 1 <= j) xi.y = 0.;
                }
                if ((i >> 1) & 1) {
                    d0 = fma(xi.x, col[i], d0);
                    d1 = fma(xi.y, col[i + 1], d1);
                } else {
                    d2 = fma(xi.x, col[i], d2);
                    d3 = fma(xi.y, col[i + 1], d3);
                }
            }
            double cj = 0.;
#pragma unroll
            for (int i = JB; i < JE; ++i)
                if (i == j) cj = col[i];
            const double scal = S.scal[j];
            const double w = S.tau[j] * fma((d0 + d1) + (d2 + d3), scal, cj);
            cj -= w;
            const double ws = -(w * scal);
#pragma unroll
            for (int i = RB; i < N; i += 2) {
                double2 xi = *reinterpret_cast<const double2 *>(x + i);
                if (i < JE) {
                    if (i <= j) xi.x = 0.;
                    if (i + 1 <= j) xi.y = 0.;
                }
                col[i] = fma(ws, xi.x, col[i]);
                col[i + 1] = fma(ws, xi.y, col[i + 1]);
            }
#pragma unroll
            for (int i = JB; i < JE; ++i)
                if (i == j) col[i] = cj;
        }
    }
}

template <int N>
__device__ __forceinline__ void qr_rolled(const Mem<N> &S, int c, bool live, double (&col)[N], double *Rout)
{
    static_assert(N % 2 == 0, "rows are processed in aligned pairs");
    constexpr int B1 = (N / 3) & ~1, B2 = (2 * N / 3) & ~1;
    const int m = S.m;
    __syncthreads();                       // every thread is done with its private facc column
    qr_factor_group<N, 0, B1>(S, c, live, col, Rout);
    qr_factor_group<N, B1, B2>(S, c, live, col, Rout);
    qr_factor_group<N, B2, N>(S, c, live, col, Rout);
    if (Rout != nullptr && live)
        for (int i = c + 1; i < m; ++i) Rout[i * m + c] = 0.;   // strictly lower part of column c
    __syncthreads();                       // the last reflector, tau and scal are published
#pragma unroll
    for (int i = 0; i < N; ++i) col[i] = i == c ? 1. : 0.;
    qr_formq_group<N, B2, N>(S, c, live, col);
    qr_formq_group<N, B1, B2>(S, c, live, col);
    qr_formq_group<N, 0, B1>(S, c, live, col);
    if (live) {
        double *fmc = S.fm + c;
#pragma unroll
        for (int i = 0; i < N; ++i) fmc[i * m] = col[i];
    }
}

// ---- Cholesky QR on the FP64 tensor cores (steps whose Q and R nobody looks at) -----------------------------------------
// The Benettin loop needs q, r = qr(prop @ q) every step (lyapunov.py:602-604), but between two records only
// log|diag r| (:611) and the SUBSPACES spanned by the leading columns of q are used: Householder's Q does not depend on
// the signs of the columns of the matrix it factorises (a reflector is the same for x and -x), so a step may return
// Q D with any diagonal D = +-1 without changing the Q, R of a later Householder step.  For those steps the
// factorisation is done as  G = A^T A  (a small GEMM per member: mma.sync.m8n8k4.f64, one warp per member, the
// fragments straight from the row-major A the tangent step left in shared memory),  G = R^T R  (right-looking
// Cholesky, lane = column in registers, one warp-level hand-over per pivot instead of a block barrier per reflector),
// Q = A R^-1  (forward substitution, two rows per thread).  The chain of dependent instructions per column is
// shuffle + rsqrt + multiply + FMA instead of Householder's reduction + sqrt + two reciprocals + update, and the Gram
// matrix -- half of the O(n^3) work -- runs on the tensor pipe with 1/8 of the instructions.  Orthogonality of Cholesky
// QR degrades with cond(A)^2: A is an orthonormal basis propagated over ONE step, cond(A) ~ exp((l_1 - l_n) dt) = O(1);
// a pivot below CHOL_PIVOT_MIN of the largest squared column norm (cond^2 > 1 / CHOL_PIVOT_MIN, or a rank-deficient
// basis) sends THAT MEMBER through the Householder code for the step (the others keep their result: a member's numbers
// must not depend on who shares its block).  Steps whose Q or R is recorded or returned (vector records, the Ginelli
// pass, the start and the final basis) always take Householder, so recorded vectors keep np.linalg.qr's signs.
constexpr double CHOL_PIVOT_MIN = 1e-4;
#ifdef CHOL_PROF
__device__ long long chol_prof[4];
#endif

// compile-time loop: nvcc stops honouring `#pragma unroll` on the outer loop of a large triangular nest (it unrolls by
// four and indexes the register arrays dynamically, i.e. puts them in local memory)
template <int I, int E, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, E>(f);
    }
}

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 1 / sqrt(x) for normal positive x: hardware seed (2^-22) + two Newton steps (2^-43, then to rounding)
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h, y * y, 0.5);
    y = fma(y, e, y);
    e = fma(-h, y * y, 0.5);
    return fma(y, e, y);
}

// x[k] -= q * row[k] for k = J + 1 .. N - 1 (row 16-byte aligned at even k, N even): LDS.128 where a pair is aligned
template <int N, int J>
__device__ __forceinline__ void row_axpy(double (&x)[N], const double *row, double q)
{
    constexpr int K0 = (J + 2) & ~1;                  // first even k > J
    if (K0 != J + 1) x[J + 1] = fma(-q, row[J + 1], x[J + 1]);
#pragma unroll
    for (int k = K0; k < N; k += 2) {
        const double2 r = *reinterpret_cast<const double2 *>(row + k);
        x[k] = fma(-q, r.x, x[k]);
        x[k + 1] = fma(-q, r.y, x[k + 1]);
    }
}

// the same for two vectors sharing the loads of the row
template <int N, int J>
__device__ __forceinline__ void row_axpy2(double (&x)[N], double (&y)[N], const double *row, double q, double p)
{
    constexpr int K0 = (J + 2) & ~1;
    if (K0 != J + 1) {
        x[J + 1] = fma(-q, row[J + 1], x[J + 1]);
        y[J + 1] = fma(-p, row[J + 1], y[J + 1]);
    }
#pragma unroll
    for (int k = K0; k < N; k += 2) {
        const double2 r = *reinterpret_cast<const double2 *>(row + k);
        x[k] = fma(-q, r.x, x[k]);
        x[k + 1] = fma(-q, r.y, x[k + 1]);
        y[k] = fma(-p, r.x, y[k]);
        y[k + 1] = fma(-p, r.y, y[k + 1]);
    }
}

// ONE WARP, one member: S.fm holds A (N x m, row-major, ld = m).  MP >= m is the COMPILE-TIME column capacity: R
// (rows < m) is left row-major with leading dimension MP and columns >= m zero in S.facc, so that the inner loops
// carry no run-time bounds; 1 / R_jj in S.scal, R_jj in S.rdiag; the verdict of the pivot test is returned and left for
// chol_passed (S.fm is not touched either way).
// FULLM: m == MP == N, no run-time bound anywhere -- the pivot steps then form one basic block, and ptxas overlaps the
// rsqrt chain of pivot j + 1 with the trailing update of pivot j.
// Columns in registers: lane l holds column l + E of G (rows <= l + E), E = max(MP - 32, 0); the first E columns have
// at most E rows above the diagonal and ride along in lanes 0 .. E - 1 as E extra values.
template <int N, int MP, bool FULLM>
__device__ __forceinline__ bool chol_factor(const Mem<N> &S, int lane)
{
    static_assert(MP % 2 == 0 && MP <= 48 && MP <= N && (!FULLM || MP == N), "pairs of columns; MP - 32 short extra columns");
    constexpr int MT = (MP + 7) / 8, KS = (N + 3) / 4;
    constexpr int E = MP > 32 ? MP - 32 : 0, EA = E > 0 ? E : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int m = FULLM ? N : S.m;
    const int ar = lane >> 2, ac = lane & 3;
    const double *A = S.fm;
    double *G = S.facc;
#ifdef CHOL_PROF
    long long tp0 = clock64();
#endif
    {
        // G = A^T A: the A operand of tile (tm, tn) is (A^T)[8 tm + ar][4 ks + ac], the B operand A[4 ks + ac][8 tn + ar]
        // -- the same fragment shape, so one load per column tile and k step serves both
        // Few tiles (partial bases): the k steps of a tile are one chain of dependent DMMAs and nothing else is in
        // flight, so the steps alternate between two accumulator sets and the loop is unrolled.
        constexpr int SETS = MT <= 3 ? 2 : 1;
        double acc[SETS][MT][MT][2];
#pragma unroll
        for (int q = 0; q < SETS; ++q)
#pragma unroll
            for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                for (int tn = 0; tn < MT; ++tn) acc[q][tm][tn][0] = acc[q][tm][tn][1] = 0.;
        auto kstep = [&](int ks, auto qc) {
            constexpr int q = decltype(qc)::value;
            const int row = 4 * ks + ac;
            double f[MT];
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                const int col = 8 * t + ar;
                f[t] = (row < N && col < m) ? A[row * m + col] : 0.;
            }
#pragma unroll
            for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                for (int tn = tm; tn < MT; ++tn) dmma_884(acc[q][tm][tn][0], acc[q][tm][tn][1], f[tm], f[tn]);
        };
        if (SETS == 1) {
#pragma unroll 1
            for (int ks = 0; ks < KS; ++ks) kstep(ks, std::integral_constant<int, 0>{});
        } else {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (ks & 1)
                    kstep(ks, std::integral_constant<int, SETS - 1>{});
                else
                    kstep(ks, std::integral_constant<int, 0>{});
            }
        }
        __syncwarp();
#pragma unroll
        for (int tm = 0; tm < MT; ++tm)
#pragma unroll
            for (int tn = tm; tn < MT; ++tn) {
                const int r = 8 * tm + ar, c0 = 8 * tn + 2 * ac;
                if (r < m && c0 < MP)
                    *reinterpret_cast<double2 *>(G + r * MP + c0) =
                        SETS == 2 ? make_double2(acc[0][tm][tn][0] + acc[SETS - 1][tm][tn][0],
                                                 acc[0][tm][tn][1] + acc[SETS - 1][tm][tn][1])
                                  : make_double2(acc[0][tm][tn][0], acc[0][tm][tn][1]);
            }
        __syncwarp();
    }
#ifdef CHOL_PROF
    long long tp1 = clock64();
#endif
    const int cm = lane + E;                           // the lane's main column
    const bool hasm = cm < MP, hase = lane < E;
    double gc[MP], ge[EA];
#pragma unroll
    for (int i = 0; i < MP; ++i) gc[i] = (i < m && i <= cm && hasm) ? G[i * MP + cm] : 0.;
#pragma unroll
    for (int i = 0; i < EA; ++i) ge[i] = (i < m && i <= lane && hase) ? G[i * MP + lane] : 0.;
    double gmax = lane < m ? G[lane * MP + lane] : 0.;
    if (32 + lane < m) gmax = fmax(gmax, G[(32 + lane) * MP + 32 + lane]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(FULL, gmax, o));
    const double pmin = CHOL_PIVOT_MIN * gmax;
    bool ok = gmax > 0.;
    double p = __shfl_sync(FULL, E > 0 ? ge[0] : gc[0], 0);
    double rinv = fast_rsqrt(p);
    __syncwarp();
#ifdef CHOL_PROF
    long long tp2 = clock64();
#endif
    static_for<0, MP>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if (FULLM || j < m) {                          // uniform
            ok = ok && (p > pmin);
            const double rjc = gc[j] * rinv;
            const double rje = j < E ? ge[j < E ? j : 0] * rinv : 0.;
            if (lane == 0) {
                S.rdiag[j] = p * rinv;
                S.scal[j] = rinv;
            }
            // the next pivot straight from its owner's registers: G_{j+1,j+1} - R_{j,j+1}^2 needs only the owner's own
            // R_{j,j+1}, not the round trip of row j through shared memory
            if (j + 1 < MP) {
                constexpr int j1 = j + 1 < MP ? j + 1 : 0;
                const double pn = j1 < E ? fma(-rje, rje, ge[j1 < E ? j1 : 0]) : fma(-rjc, rjc, gc[j1]);
                p = __shfl_sync(FULL, pn, j1 < E ? j1 : j1 - E);
            }
            double *row = G + j * MP;
            if (hasm && cm >= j) row[cm] = rjc;
            if (j < E && hase && lane >= j) row[lane] = rje;
            __syncwarp();
            if (j + 1 < MP) rinv = fast_rsqrt(p);
            row_axpy<MP, j>(gc, row, rjc);
            if (j + 1 < E) {
#pragma unroll
                for (int i = j + 1; i < E; ++i) ge[i] = fma(-row[i], rje, ge[i]);
            }
        }
    });
#ifdef CHOL_PROF
    if (lane == 0 && threadIdx.x == 0 && blockIdx.x == 0) {
        long long tp3 = clock64();
        chol_prof[0] += tp1 - tp0; chol_prof[1] += tp2 - tp1; chol_prof[2] += tp3 - tp2;
    }
#endif
    if (lane == 0) S.tau[0] = ok ? 1. : 0.;            // the member's verdict (chol_passed)
    return ok;
}

template <int N>
__device__ __forceinline__ bool chol_passed(const Mem<N> &S) { return S.tau[0] != 0.; }

// Q = A R^-1 by forward substitution, TWO rows of one member per thread (the loads of R's rows are broadcasts that
// cost the load/store unit as much as any other load: two rows per load halve them), A (S.fm) is overwritten.
// idx enumerates (member, row pair); rows r and r + N / 2.
template <int N, int MP, bool FULLM>
__device__ __forceinline__ void chol_solve(const Mem<N> &S, int pr)
{
    const int m = FULLM ? N : S.m;               // m <= MP
    const double *R = S.facc;
    double *arow = S.fm + pr * m, *brow = S.fm + (pr + N / 2) * m;
    double a[MP], b[MP];
    if (FULLM) {                                       // m = N is even: rows are 16-byte aligned
#pragma unroll
        for (int k = 0; k < MP; k += 2) {
            const double2 va = *reinterpret_cast<const double2 *>(arow + k), vb = *reinterpret_cast<const double2 *>(brow + k);
            a[k] = va.x, a[k + 1] = va.y, b[k] = vb.x, b[k + 1] = vb.y;
        }
    } else {
        // columns m .. MP - 1 do not exist (they would be the next row, which another thread is rewriting)
#pragma unroll
        for (int k = 0; k < MP; ++k) {
            a[k] = k < m ? arow[k] : 0.;
            b[k] = k < m ? brow[k] : 0.;
        }
    }
    static_for<0, MP>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if (FULLM || j < m) {                          // uniform
            const double s = S.scal[j];
            const double qa = a[j] * s, qb = b[j] * s;
            a[j] = qa;
            b[j] = qb;
            if (j + 1 < MP) row_axpy2<MP, j>(a, b, R + j * MP, qa, qb);
        }
    });
    if (FULLM) {
#pragma unroll
        for (int k = 0; k < MP; k += 2) {
            *reinterpret_cast<double2 *>(arow + k) = make_double2(a[k], a[k + 1]);
            *reinterpret_cast<double2 *>(brow + k) = make_double2(b[k], b[k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < MP; ++k)
            if (k < m) {
                arow[k] = a[k];
                brow[k] = b[k];
            }
    }
}

// the block's part of one Cholesky-QR step: A in S.fm of every member -> Q in S.fm, R_jj in S.rdiag, for the members
// whose factorisation passed the pivot test (chol_passed); the others keep A.  The verdict is per MEMBER, so that a
// member's results do not depend on who shares its block (members sharded over devices or cut into batches must stay
// bitwise what one launch gives).  Returns true when some member of the block refused.  Contains block barriers.
template <int N, int MP, bool FULLM>
__device__ __forceinline__ bool cholqr_block(double *smem, int stride, int jv, int m, int G, long n_members)
{
    const int t = threadIdx.x;
    __syncthreads();                                   // A = the columns tangent_step left in S.fm
    bool ok = true;
    for (int gg = t >> 5; gg < G; gg += blockDim.x >> 5)
        if ((long)blockIdx.x * G + gg < n_members)
            ok = chol_factor<N, MP, FULLM>(carve<N>(smem + (size_t)gg * stride, jv, m), t & 31) && ok;
    const bool refused = __syncthreads_or(!ok);
    // member index fastest: the rows of one lane are a whole row stride apart (8-way bank conflicts when neighbouring
    // lanes take neighbouring rows), the members' areas are staggered by 16 bytes modulo 128
    for (int idx = t; idx < G * (N / 2); idx += blockDim.x) {
        const int pr = idx / G, gg = idx - pr * G;
        if ((long)blockIdx.x * G + gg < n_members) {
            const Mem<N> Sg = carve<N>(smem + (size_t)gg * stride, jv, m);
            if (!refused || chol_passed<N>(Sg)) chol_solve<N, MP, FULLM>(Sg, pr);
        }
    }
    __syncthreads();
    return refused;
}

// column capacities the Cholesky QR is instantiated for: the full basis (no run-time bounds at all) and 16 / 24 / 32
// columns for partial bases (a 10-vector basis pays for 16 columns, not for N)
template <int N>
__host__ __device__ constexpr bool chol_capacity_exists(int m)
{
    return m == N || (!QGSB_PACK_LEAN && m < N && m <= 32 && (m <= 16 ? 16 : m <= 24 ? 24 : 32) <= N);
}

template <int N>
__device__ __forceinline__ bool cholqr_dispatch(double *smem, int stride, int jv, int m, int G, long n_members)
{
    if (m == N) return cholqr_block<N, N, true>(smem, stride, jv, m, G, n_members);
    if constexpr (N >= 16 && !QGSB_PACK_LEAN)
        if (m <= 16) return cholqr_block<N, 16, false>(smem, stride, jv, m, G, n_members);
    if constexpr (N >= 24 && !QGSB_PACK_LEAN)
        if (m <= 24) return cholqr_block<N, 24, false>(smem, stride, jv, m, G, n_members);
    if constexpr (N >= 32 && !QGSB_PACK_LEAN)
        if (m <= 32) return cholqr_block<N, 32, false>(smem, stride, jv, m, G, n_members);
    return true;            // not reached: the launch only asks for capacities that exist (chol_capacity_exists)
}

template <int N, class Prod>
__device__ __forceinline__ void init_member(const Mem<N> &S, int c, bool live)
{
    if (!live) return;
    const int m = S.m;
    if (Prod::kZeroFill)
        for (int q = c; q < Prod::JV; q += m) S.jv[q] = 0.;
    for (int r = c; r < N; r += m) {
        S.kst[r] = 0.;
        S.yacc[r] = 0.;
    }
    if (c == 0) {
        S.xs[0] = 1.;
        S.xs2[0] = 1.;
    }
}

// ---- plain tangent-linear integration (integrate.py:555-614) ----------------------------------------------------------
template <int N, class Prod>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
tgls_kernel(TensorView T, const __grid_constant__ TgParams P, const PackTables tab_g, int G, int stride)
{
    extern __shared__ __align__(16) double smem_pack[];
    const ShTab tab = stage_tables(tab_g, smem_pack + (size_t)G * stride, N);
    const int m = P.m, t = threadIdx.x;
    const int g = t / m, c = t - g * m;
    const long member = (long)blockIdx.x * G + g;
    const bool live = g < G && member < P.n_members;
    const Mem<N> S = carve<N>(smem_pack + (size_t)(live ? g : 0) * stride, Prod::JV, m);
    // the tendencies are dealt out member-fastest (see tangent_step)
    const int cq = t / G, gq = t - cq * G;
    const bool liveq = cq < m && (long)blockIdx.x * G + gq < P.n_members;
    const Mem<N> Sq = carve<N>(smem_pack + (size_t)(liveq ? gq : 0) * stride, Prod::JV, m);
    const int nm = N * m;
    double col[N];
    init_member<N, Prod>(S, c, live);
    if (live)
        for (int r = c; r < N; r += m) S.y[r] = P.y[member * N + r];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        col[i] = live ? P.fm[member * nm + i * m + c] : 0.;
        if (live) S.fm[i * m + c] = col[i];
    }
    __syncthreads();
    long iw = 0;
#pragma unroll 1
    for (long ti = 0; ti < P.n_steps; ++ti) {
        if (P.rec_y && P.write_steps > 0 && ti % P.write_steps == 0) {
            if (live) {
                double *ry = P.rec_y + ((size_t)iw * P.n_members + member) * N;
                double *rf = P.rec_fm + ((size_t)iw * P.n_members + member) * nm;
                for (int r = c; r < N; r += m) ry[r] = S.y[r];
#pragma unroll
                for (int i = 0; i < N; ++i) rf[i * m + c] = col[i];
            }
            ++iw;
        }
        tangent_step<N, Prod, true>(T, tab, P, S, P.dt[ti], col, c, live, Sq, cq, liveq);
        __syncthreads();
    }
    if (live) {
        if (P.rec_y) {
            double *ry = P.rec_y + ((size_t)(P.n_records - 1) * P.n_members + member) * N;
            double *rf = P.rec_fm + ((size_t)(P.n_records - 1) * P.n_members + member) * nm;
            for (int r = c; r < N; r += m) ry[r] = S.y[r];
#pragma unroll
            for (int i = 0; i < N; ++i) rf[i * m + c] = col[i];
        }
        for (int r = c; r < N; r += m) P.y[member * N + r] = S.y[r];
#pragma unroll
        for (int i = 0; i < N; ++i) P.fm[member * nm + i * m + c] = col[i];
    }
}

// ---- Benettin loop (lyapunov.py:471-632) ---------------------------------------------------------------------------------
template <int N, class Prod, bool ROLLED>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
lyap_kernel(TensorView T, const __grid_constant__ TgParams P, const PackTables tab_g, int G, int stride, int qr_flags)
{
    extern __shared__ __align__(16) double smem_pack[];
    const ShTab tab = stage_tables(tab_g, smem_pack + (size_t)G * stride, N);
    const int m = P.m, t = threadIdx.x;
    const int g = t / m, c = t - g * m;
    const long member = (long)blockIdx.x * G + g;
    const bool live = g < G && member < P.n_members;
    const Mem<N> S = carve<N>(smem_pack + (size_t)(live ? g : 0) * stride, Prod::JV, m);
    const int nm = N * m;
    const long steps = P.n_pre + P.n_rec;
    const long R = P.n_records;
    // For the QR the columns are dealt out a second way, member index fastest: a warp then holds the same few
    // columns of all the members, so the warps whose columns are already finished skip a reflector altogether
    // (with the consecutive layout every warp has live columns until the very end).  The tangent steps keep the
    // consecutive layout, whose broadcast reads of J hit one or two addresses per warp.
    const bool remap = (qr_flags & 1) != 0, chol = (qr_flags & 2) != 0;
    const int cq = remap ? t / G : c, gq = remap ? t - cq * G : g;
    const long memberq = (long)blockIdx.x * G + gq;
    const bool liveq = remap ? (cq < m && memberq < P.n_members) : live;
    const Mem<N> Sq = carve<N>(smem_pack + (size_t)(liveq ? gq : 0) * stride, Prod::JV, m);
    double col[N];
    init_member<N, Prod>(S, c, live);
    // rows of the state (y, Y): dealt out like the tendencies (Sq, cq -- see tangent_step), here and everywhere below
    if (liveq)
        for (int r = cq; r < N; r += m) Sq.Y[r] = P.y[memberq * N + r];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        col[i] = live ? P.fm[member * nm + i * m + c] : 0.;
        if (live) S.fm[i * m + c] = col[i];
    }
    if (live) S.rdiag[c] = P.r0 ? P.r0[((size_t)member * m + c) * m + c] : 0.;
    __syncthreads();
    const size_t sbase = (P.stored && liveq) ? tile_base(memberq, N) : 0;
    long iw = 0;
    double mexp = 0.;
    // step -1 (P.qr_at_start): the start matrix was drawn on the device and is only factorised here -- the
    // reference's q, r = qr(random((n_dim, n_vec))) of lyapunov.py:592-593 -- without duplicating the QR code
#pragma unroll 1
    for (long step = P.qr_at_start ? -1 : 0; step < steps; ++step) {
        const bool real = step >= 0;
        if (real && P.stored) {                                           // lyapunov.py:513 / :527
            if (liveq) {
                const double *src = P.stored + (size_t)P.start_idx[step] * N * P.stored_ld + sbase;
                for (int r = cq; r < N; r += m) Sq.Y[r] = src[(size_t)r * TILE];
            }
            __syncthreads();
        }
        if (real && step >= P.n_pre) {
            const long ti = step - P.n_pre;
            // :611 / :531 -- only where it is written: at a record, and in the last step for the record after the loop
            const bool rec_now = P.write_steps > 0 && ti % P.write_steps == 0;
            if (live && (rec_now || step + 1 == steps)) mexp = log(fabs(S.rdiag[c])) / P.dt_macro[step];
            if (P.q_all && live) {
                double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + ti) * nm;
#pragma unroll
                for (int i = 0; i < N; ++i) qa[i * m + c] = col[i];
            }
            if (rec_now) {
                if (liveq) {
                    const long rc = P.forward == 1 ? R - 1 - iw : iw;
                    double *ry = P.rec_y + ((size_t)rc * P.n_members + memberq) * N;
                    for (int r = cq; r < N; r += m) ry[r] = Sq.Y[r];
                }
                if (live) {
                    const long rc = P.forward == 1 ? R - 1 - iw : iw;
                    double *rv = P.rec_fm + ((size_t)rc * P.n_members + member) * nm;
                    double *re = P.rec_exp + ((size_t)rc * P.n_members + member) * m;
                    if (P.rec_fm) {
#pragma unroll
                        for (int i = 0; i < N; ++i) rv[i * m + c] = col[i];
                    }
                    re[c] = mexp;
                }
                ++iw;
            }
        }
        // propagate the basis over the micro steps starting from the stored point (:598-600)
        long q0 = 0, q1 = 0;
        if (real) {
            if (liveq)
                for (int r = cq; r < N; r += m) Sq.y[r] = Sq.Y[r];
            q0 = P.sub_ptr[step];
            q1 = P.sub_ptr[step + 1];
            for (long q = q0; q < q1; ++q) tangent_step<N, Prod, (N <= 24)>(T, tab, P, S, P.sub_dt[q], col, c, live, Sq, cq, liveq);
        }
        // q, r = qr(prop @ q)   (:602-604)
        // hh / hhq: the thread's Householder roles (its own column; with the remap the column it factorises)
        bool householder = true, partial = false, hh = live, hhq = liveq;
        if (chol && real && step + 1 < steps && !P.q_all && !P.r_all &&
            !(P.rec_fm && P.write_steps > 0 && step + 1 >= P.n_pre && (step + 1 - P.n_pre) % P.write_steps == 0)) {
            // nobody sees this step's Q or R: Cholesky QR (see chol_factor); members that refuse a pivot, and only
            // they, go through the Householder code below
            householder = partial = cholqr_dispatch<N>(smem_pack, stride, Prod::JV, m, G, P.n_members);
            if (partial) {
                hh = live && !chol_passed<N>(S);
                hhq = liveq && !chol_passed<N>(Sq);
            }
            // Q -- or, for a member that refused, the untouched A: the column does not stay in registers across the
            // factorisation (72 registers that the Cholesky phase needs)
            if (live) {
#pragma unroll
                for (int i = 0; i < N; ++i) col[i] = S.fm[i * m + c];
            }
        }
        if (householder) {
            double *Rout = (real && P.r_all && liveq && step >= P.r_first)
                               ? P.r_all + ((size_t)memberq * (steps - P.r_first) + (step - P.r_first)) * m * m : nullptr;
            if (remap) {                     // hand the columns over through the fm area (tangent_step left them there)
                __syncthreads();
                if (hhq) {
#pragma unroll
                    for (int i = 0; i < N; ++i) col[i] = Sq.fm[i * m + cq];
                }
            }
            if (ROLLED)
                qr_rolled<N>(Sq, cq, hhq, col, Rout);
            else
                qr<N>(Sq, cq, hhq, col, Rout);
            if (remap || partial) {          // (the factorisation leaves unit vectors in the registers of idle threads)
                __syncthreads();
                if (live) {
#pragma unroll
                    for (int i = 0; i < N; ++i) col[i] = S.fm[i * m + c];
                }
            }
        }
        if (!real) continue;
        if (P.forward == 2 || (!P.stored && q1 - q0 == 1 && P.sub_dt[q0] == P.dt_macro[step])) {
            // Ginelli forward pass follows the micro steps; and with a single micro step of the macro length the
            // "stored" trajectory point (:601 / :622) is bit-for-bit the state the tangent step just produced
            if (liveq)
                for (int r = cq; r < N; r += m) Sq.Y[r] = Sq.y[r];
        } else if (!P.stored) {                           // next stored-trajectory point (:601 / :622)
            nl_step<N>(T, tab, P, Sq, P.dt_macro[step], cq, liveq);
        }
    }
    if (liveq) {
        if (P.stored) {
            const double *src = P.stored + (size_t)P.final_idx * N * P.stored_ld + sbase;
            for (int r = cq; r < N; r += m) Sq.Y[r] = src[(size_t)r * TILE];
        }
        const long rc = P.forward == 1 ? 0 : R - 1;                       // :628-630 / :548-550
        double *ry = P.rec_y + ((size_t)rc * P.n_members + memberq) * N;
        for (int r = cq; r < N; r += m) {
            ry[r] = Sq.Y[r];
            P.y[memberq * N + r] = Sq.Y[r];
        }
    }
    if (live) {
        const long rc = P.forward == 1 ? 0 : R - 1;
        double *rv = P.rec_fm + ((size_t)rc * P.n_members + member) * nm;
        double *re = P.rec_exp + ((size_t)rc * P.n_members + member) * m;
        if (P.rec_fm) {
#pragma unroll
            for (int i = 0; i < N; ++i) rv[i * m + c] = col[i];
        }
        re[c] = mexp;
        if (P.q_all) {
            double *qa = P.q_all + ((size_t)member * (P.n_rec + 1) + P.n_rec) * nm;
#pragma unroll
            for (int i = 0; i < N; ++i) qa[i * m + c] = col[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) P.fm[member * nm + i * m + c] = col[i];
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
// table_bytes: size of the ELL tables, which are always staged in shared memory behind the members' areas
// n_members / sm_count > 0: among the block sizes that fit, take the one that minimises (waves of blocks) x (time of a
// block, modelled as half fixed and half proportional to its members): 8192 members with 10 vectors are 328 blocks of
// 25 = 2.2 waves that cost three; 432 blocks of 19 fill three waves with faster blocks.  A launch that does not fill the
// SMs gets small blocks for the same reason.
template <int N>
inline Geometry geometry(int jv, int m, size_t smem_limit, size_t table_bytes, long n_members = 0, int sm_count = 0)
{
    Geometry geo;
    const Carve<N> c(jv, m);
    geo.jv = jv;
    geo.stride = c.total();
    const size_t per_member = (size_t)geo.stride * sizeof(double);
    int G = MAX_THREADS / m;
    if (const char *env = getenv("QGSB_PACK_G")) G = std::max(1, std::min(G, atoi(env)));   // A/B measurements
    if (table_bytes >= smem_limit) G = 0;
    else if ((size_t)G * per_member + table_bytes > smem_limit) G = (int)((smem_limit - table_bytes) / per_member);
    if (G > 1 && n_members > 0 && sm_count > 0 && !getenv("QGSB_PACK_G")) {
        int best = G;
        long best_cost = -1;
        for (int g = G; g >= (G + 1) / 2; --g) {
            const long blocks = (n_members + g - 1) / g, waves = (blocks + sm_count - 1) / sm_count;
            const long cost = waves * (2 * g + G);         // a block's time: about half fixed, half per member
            if (best_cost < 0 || cost < best_cost) best = g, best_cost = cost;
        }
        G = best;
    }
    geo.G = G;
    geo.threads = G > 0 ? ((G * m + 31) / 32) * 32 : 0;
    geo.smem = (size_t)G * per_member + table_bytes;
    return geo;
}

inline size_t table_bytes(const PackTables &tab, int n)
{
    size_t b = 0;
    if (tab.f_ent) b += (size_t)tab.EF * n * sizeof(PEnt);
    if (tab.j_ent) b += (size_t)tab.EJ * tab.npos * sizeof(PEnt);
    return b;
}

// launches the packed kernel for one policy pair; returns cudaErrorInvalidValue when it does not fit
template <int N, class Fwd, class Adj>
inline cudaError_t launch(const TensorView &T, const TgParams &P, const PackTables &tables, bool lyap,
                          size_t smem_limit, cudaStream_t stream)
{
    static_assert(Fwd::JV == Adj::JV, "both directions of a product share one Jacobian layout");
    if (P.m < 1 || P.m > MAX_THREADS) return cudaErrorInvalidValue;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const Geometry geo = geometry<N>(Fwd::JV, P.m, smem_limit, table_bytes(tables, N), P.n_members, sms);
    if (geo.G < 1) return cudaErrorInvalidValue;
    const PackTables &tab = tables;
    const unsigned blocks = (unsigned)((P.n_members + geo.G - 1) / geo.G);
    auto go = [&](auto kernel) -> cudaError_t {
        if (geo.smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem);
            if (e != cudaSuccess) return e;
        }
        kernel<<<blocks, geo.threads, geo.smem, stream>>>(T, P, tab, geo.G, geo.stride);
        return cudaGetLastError();
    };
    auto go_lyap = [&](auto kernel) -> cudaError_t {
        if (geo.smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem);
            if (e != cudaSuccess) return e;
        }
        const char *env = getenv("QGSB_QR_REMAP");
        const int remap = env ? (env[0] != '0') : 1;
        // Cholesky QR for the steps whose Q, R are not observed (chol_factor); QGSB_QR_CHOL=0 keeps Householder everywhere
        // [B200, 8192 members, exponents only, against Householder everywhere: MAOOAM-36 36 vectors x1.51, 20 vectors
        // x1.42, 10 vectors x1.04, 5 vectors x0.98; RP-20 x1.19; dynamic-T 38 vectors x1.23]: on from 8 vectors, where a
        // column capacity exists (chol_capacity_exists)
        const char *envc = getenv("QGSB_QR_CHOL");
        const int chol = chol_capacity_exists<N>(P.m) && (envc ? (envc[0] != '0') : (P.m >= 8));
        kernel<<<blocks, geo.threads, geo.smem, stream>>>(T, P, tab, geo.G, geo.stride, remap | (chol << 1));
        return cudaGetLastError();
    };
    if (lyap) {
        // few vectors: the rolled factorisation (one resident loop body instead of n_vec unrolled reflectors) is faster
        // [B200: MAOOAM-36, 10 vectors +19 %; 36 vectors -8 %].
        // Three other forms were measured in round 2 and removed (profiles/r02_benettin_qr_pipelined_ab.log,
        // r02_benettin_split_ab.log, r02_benettin_qr_panel2_ab.log): reflectors handed over through progress flags
        // instead of block barriers (-16 %: the factorisation is ONE dependent chain per block, the other warps have
        // nothing to overlap it with); the step as two launches, propagation + stand-alone QR with two blocks per SM
        // (-15 %: 26 % / 33 % of the two kernels go into moving the state through L2, and at 128 registers the
        // factorisation spills); two reflectors per barrier with the second one applied algebraically (-5 % at 36
        // vectors, +5 % at 20: half the barriers, but the chain of dependent scalar instructions -- square root, two
        // reciprocals, ~19 deep per reflector at 12-19 cycles each -- is what a reflector costs, and it stays).
        const char *env = getenv("QGSB_QR_ROLLED");               // 0 / 1 forces one of them (A/B measurements)
        const bool rolled = QGSB_PACK_LEAN || (env ? (env[0] != '0') : (3 * P.m <= N));
        if (rolled) return P.adjoint ? go_lyap(lyap_kernel<N, Adj, true>) : go_lyap(lyap_kernel<N, Fwd, true>);
        if constexpr (!QGSB_PACK_LEAN)
            return P.adjoint ? go_lyap(lyap_kernel<N, Adj, false>) : go_lyap(lyap_kernel<N, Fwd, false>);
        return cudaErrorInvalidValue;       // not reached
    }
    return P.adjoint ? go(tgls_kernel<N, Adj>) : go(tgls_kernel<N, Fwd>);
}

}  // namespace pack
}  // namespace qgsb
