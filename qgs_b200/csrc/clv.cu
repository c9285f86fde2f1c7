// clv.cu -- covariant Lyapunov vectors by the method of Ginelli et al., entirely on the device.
//
// Replaces _compute_clv_gin_jit (qgs/toolbox/lyapunov.py:1174-1288) and the helpers it calls,
// solve_triangular_matrix / normalize_matrix_columns (qgs/functions/util.py:56-98):
//
//   parts 1-3  forward Benettin pass from t0 to tc that stores the basis Q and the state at every step of
//              [ta, tb] and the factor R of every step of [ta, tc]        -> the Benettin kernels (mode 2);
//   parts 4-5  backward recursion A <- normalise(R_ti^{-1} A) from tc to ta, recording Q_ti A, the local
//              exponents -log|norm| / dt and the state on the way from tb to ta  -> ginelli_kernel below.
//
// The reference keeps R and Q of every step of one trajectory in the worker's memory; here they stay in HBM
// (about 2 * 8 n m bytes per member and step) and only the records come back, so the members are processed in
// batches sized to a memory budget.  One block per member runs the backward recursion; thread c owns column
// c of the upper-triangular m x m matrix A (shared memory), R_ti is staged in shared memory and read as
// broadcasts, the triangular solve is the column-oriented back substitution that LAPACK's dgetrs performs for
// the reference's np.linalg.solve on a triangular block.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "tgls_shared.cuh"

namespace qgsb {

struct GinelliParams {
    long n_members;          // members of this batch
    int n, m;
    long tw, tew;            // steps of [ta, tb] and of [ta, tc]
    long write_steps, n_records;
    double noise_pert;
    const double *r_all;     // (Nb, tew, m, m)
    const double *rec_y;     // (tew + 1, Nb, n)     state at every step (write_steps = 1 records of the forward pass)
    const double *rec_q;     // (tew + 1, Nb, n, m)  basis at every step
    const double *am0;       // (Nb, m, m) start matrices (upper triangular, unit columns)
    const double *noise;     // (Nb, tew, m) or null
    const double *dte;       // (tw + 1)
    double *out_y;           // (Nb, n, R)
    double *out_exp;         // (Nb, m, R)
    double *out_vec;         // (Nb, n, m, R)
};

__global__ void __launch_bounds__(1024) ginelli_kernel(const __grid_constant__ GinelliParams P)
{
    extern __shared__ __align__(16) double smem_clv[];
    const int n = P.n, m = P.m, c = threadIdx.x;
    const long member = blockIdx.x, R = P.n_records;
    const int lda = m + 1;
    double *Rs = smem_clv;                 // (m, m) row-major
    double *A = Rs + (size_t)m * m;        // column c at A + c * lda
    const bool act = c < m;
    double *bcol = A + (size_t)c * lda;
    if (act) {
        const double *a0 = P.am0 + (size_t)member * m * m;
        for (int i = 0; i < m; ++i) bcol[i] = i <= c ? a0[(size_t)i * m + c] : 0.;
    }
    long iw = 1;
    double mloc = 1.;
    for (long ti = P.tew - 1; ti >= 0; --ti) {
        __syncthreads();
        const double *Rg = P.r_all + ((size_t)member * P.tew + ti) * m * m;
        for (int q = threadIdx.x; q < m * m; q += blockDim.x) Rs[q] = Rg[q];
        __syncthreads();
        if (act) {
            // x[:c+1] = solve(R[:c+1, :c+1], b[:c+1])        util.py:94-97
            for (int k = m - 1; k >= 0; --k) {
                if (k <= c) {
                    const double xk = bcol[k] / Rs[(size_t)k * m + k];
                    bcol[k] = xk;
                    for (int i = 0; i < k; ++i) bcol[i] -= xk * Rs[(size_t)i * m + k];
                }
            }
            if (P.noise) bcol[c] += P.noise[((size_t)member * P.tew + ti) * m + c] * P.noise_pert;   // :1259-1262
            double s2 = 0.;
            for (int i = 0; i <= c; ++i) s2 += bcol[i] * bcol[i];
            mloc = sqrt(s2);                                   // util.py:72-73
            for (int i = 0; i <= c; ++i) bcol[i] /= mloc;
        }
        if (ti <= P.tw) {
            const bool periodic = P.write_steps > 0 && (P.tw - ti) % P.write_steps == 0;     // :1273-1277
            for (int pass = 0; pass < 2; ++pass) {
                long rec;
                if (pass == 0) {
                    if (!periodic) continue;
                    rec = R - iw;
                } else {
                    if (ti != 0) continue;                     // :1279-1281: record 0 holds the values of ti = 0
                    rec = 0;
                }
                if (rec < 0 || rec >= R) continue;
                const double *yq = P.rec_y + ((size_t)ti * P.n_members + member) * n;
                for (int i = threadIdx.x; i < n; i += blockDim.x) P.out_y[((size_t)member * n + i) * R + rec] = yq[i];
                if (act) {
                    P.out_exp[((size_t)member * m + c) * R + rec] = -log(fabs(mloc)) / P.dte[ti];
                    const double *Q = P.rec_q + ((size_t)ti * P.n_members + member) * n * m;
                    for (int i = 0; i < n; ++i) {
                        double v = 0.;
                        for (int k = 0; k <= c; ++k) v += Q[(size_t)i * m + k] * bcol[k];
                        P.out_vec[(((size_t)member * n + i) * m + c) * R + rec] = v;
                    }
                }
            }
            if (periodic) ++iw;
        }
    }
}


// ---- subspace intersection (lyapunov.py:1315-1320) ---------------------------------------------------------------
// CLV j of a record is the direction common to span(BLV_0..j) and span(FLV_0..n-1-j).  The reference takes the
// first left singular vector of BLV[:, :j+1]^T FLV[:, :n-j] (whose largest singular value is 1: the two bases are
// orthonormal and the dimensions add up to n + 1).  Equivalently u spans the null space of the j x (j+1) matrix
// C_j = FLV[:, n-j:]^T BLV[:, :j+1], a corner of the single n x n matrix P = FLV^T BLV -- so one block per
// (member, record) forms P once and finds every null vector by Gauss-Jordan elimination with pivoting inside
// the row (thread = column), no SVD.  The sign of a singular vector is arbitrary; here the component of largest
// modulus of u is positive.
struct SubspaceParams {
    long n_pairs;            // members of the batch x records
    long R;                  // records (stride of the last axis of the arrays)
    int n;
    const double *bvec;      // (Nb, n, n, R)
    const double *fvec;      // (Nb, n, n, R)
    double *vec;             // (Nb, n, n, R)
};

__global__ void __launch_bounds__(64) subspace_kernel(const __grid_constant__ SubspaceParams P)
{
    extern __shared__ __align__(16) double smem_sub[];
    const int n = P.n, tid = threadIdx.x, ld = n + 1;
    const long pair = blockIdx.x, member = pair / P.R, rec = pair % P.R;
    double *B = smem_sub;                    // (n, ld)   B[i][k] = BLV k, component i
    double *F = B + (size_t)n * ld;
    double *Pm = F + (size_t)n * ld;         // (n, ld)   Pm[a][b] = FLV_a . BLV_b
    double *C = Pm + (size_t)n * ld;         // (n, ld)   work matrix
    double *u = C + (size_t)n * ld;          // (n + 1)
    __shared__ int perm[64], s_col;
    __shared__ double s_best[2];
    __shared__ int s_arg[2];
    const size_t base = (size_t)member * n * n * P.R + rec;
    for (int q = tid; q < n * n; q += blockDim.x) {
        const int i = q / n, k = q - i * n;
        B[i * ld + k] = P.bvec[base + (size_t)q * P.R];
        F[i * ld + k] = P.fvec[base + (size_t)q * P.R];
    }
    __syncthreads();
    for (int q = tid; q < n * n; q += blockDim.x) {
        const int a = q / n, b = q - a * n;
        double acc = 0.;
        for (int i = 0; i < n; ++i) acc += F[i * ld + a] * B[i * ld + b];
        Pm[a * ld + b] = acc;
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const int cols = j + 1;              // unknowns; rows = j
        // C = Pm[n-j : n, 0 : j+1]
        for (int q = tid; q < j * cols; q += blockDim.x) {
            const int r = q / cols, c = q - r * cols;
            C[r * ld + c] = Pm[(n - j + r) * ld + c];
        }
        if (tid < cols) perm[tid] = 0;       // row + 1 once the column has been the pivot of that row
        __syncthreads();
        for (int k = 0; k < j; ++k) {
            // pivot: the largest entry of row k among the columns not used yet
            double best = -1.;
            int arg = -1;
            if (tid < cols && perm[tid] == 0) {
                best = fabs(C[k * ld + tid]);
                arg = tid;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) {
                    best = ob;
                    arg = oa;
                }
            }
            if ((tid & 31) == 0) {
                s_best[tid >> 5] = best;
                s_arg[tid >> 5] = arg;
            }
            __syncthreads();
            if (tid == 0) {
                int p = s_arg[0];
                if (s_arg[1] >= 0 && (p < 0 || s_best[1] > s_best[0])) p = s_arg[1];
                s_col = p;
            }
            __syncthreads();
            const int p = s_col;
            const double piv = C[k * ld + p];
            // Gauss-Jordan: clear column p in every other row; thread = column
            if (tid < cols && tid != p && piv != 0.) {
                const double ckc = C[k * ld + tid] / piv;
                for (int r = 0; r < j; ++r)
                    if (r != k) C[r * ld + tid] -= C[r * ld + p] * ckc;
            }
            __syncthreads();                  // everybody is done reading column p
            if (tid < j && tid != k) C[tid * ld + p] = 0.;
            if (tid == 0) perm[p] = k + 1;
            __syncthreads();
        }
        // the one column never used as pivot is the free unknown: u_free = 1, u_p = -C[row(p)][free] / C[row(p)][p]
        if (tid < cols) {
            int freec = 0;
            for (int c = 0; c < cols; ++c)
                if (perm[c] == 0) freec = c;
            double val = 1.;
            if (perm[tid] != 0) {
                const int r = perm[tid] - 1;
                const double d = C[r * ld + tid];
                val = d != 0. ? -C[r * ld + freec] / d : 0.;
            }
            u[tid] = val;
        }
        __syncthreads();
        // normalise, largest component positive
        if (tid < 32) {
            double s2 = 0., big = 0.;
            for (int c = tid; c < cols; c += 32) {
                s2 += u[c] * u[c];
                if (fabs(u[c]) > fabs(big)) big = u[c];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                const double ob = __shfl_xor_sync(0xffffffffu, big, o);
                if (fabs(ob) > fabs(big)) big = ob;
            }
            const double scale = (big < 0. ? -1. : 1.) / sqrt(s2);
            for (int c = tid; c < cols; c += 32) u[c] *= scale;
        }
        __syncthreads();
        // clv_j = BLV[:, :j+1] @ u
        if (tid < n) {
            double acc = 0.;
            for (int k = 0; k < cols; ++k) acc += B[tid * ld + k] * u[k];
            P.vec[base + ((size_t)tid * n + j) * P.R] = acc;
        }
        __syncthreads();
    }
}

}  // namespace qgsb

using namespace qgsb;

// one device's share of qgsb_clv_ginelli: members [0, N) of every member-major array, on the calling thread's device
static void ginelli_device(const qgsb_tensor *t, long N, const double *ic, int n_vec, const double *q0,
                           const double *r0, long n_pre, long n_time, long n_after, const double *dt_macro,
                           const long *sub_ptr, const double *sub_dt, int s, const double *a, const double *b,
                           long write_steps, double noise_pert, const double *am0, const double *noise,
                           const double *dte, long R, double *rec_traj, double *rec_exp, double *rec_vec,
                           double *device_ms)
{
    Context &cx = ctx();
    cudaStream_t st = cx.stream;
    const Tableau tab = make_tableau(s, a, b);
    const int n = t->view.n, m = n_vec;
    const size_t nm = (size_t)n * m, mm = (size_t)m * m;
    const long tw = n_time, tew = n_time + n_after, steps = n_pre + tew;
    {
        long L = tw + 1, r = write_steps == 0 ? 1 : (L + write_steps - 1) / write_steps;
        if (write_steps > 0 && (r - 1) * write_steps != L - 1) r += 1;
        QGSB_REQUIRE(R == r, "n_records %ld inconsistent with %ld steps / write_steps %ld", R, tw, write_steps);
    }
    const long n_sub = sub_ptr[steps];
    // members per batch from a memory budget: R factors of [ta, tc], basis and state at every step, outputs
    const size_t per_member = ((size_t)tew * mm + (size_t)(tew + 1) * (nm + n + m) + (size_t)R * (nm + n + m) + 4 * nm) * 8;
    const size_t budget = std::min<size_t>(cx.total_mem / 4, (size_t)24 << 30);
    const long batch = std::max<long>(1, std::min<long>(N, (long)(budget / std::max<size_t>(per_member, 1))));
    const long Rf = tew + 1;                                   // records of the forward pass (write_steps = 1)
    DevBuf<double> d_dtm(std::max<long>(steps, 1)), d_sub(std::max<long>(n_sub, 1)), d_dte(tw + 1);
    DevBuf<long> d_ptr(steps + 1);
    DevBuf<double> d_y((size_t)batch * n), d_q((size_t)batch * nm), d_r0, d_am((size_t)batch * mm), d_noise;
    DevBuf<double> d_rall((size_t)batch * std::max<long>(tew, 1) * mm);
    DevBuf<double> d_ry((size_t)Rf * batch * n), d_rq((size_t)Rf * batch * nm), d_re((size_t)Rf * batch * m);
    DevBuf<double> d_oy((size_t)batch * n * R), d_oe((size_t)batch * m * R), d_ov((size_t)batch * nm * R), scratch;
    if (steps) d_dtm.upload(dt_macro, steps, st);
    if (n_sub) d_sub.upload(sub_dt, n_sub, st);
    d_dte.upload(dte, tw + 1, st);
    QGSB_CUDA(cudaMemcpyAsync(d_ptr.p, sub_ptr, sizeof(long) * (steps + 1), cudaMemcpyHostToDevice, st));
    if (r0) d_r0.alloc((size_t)batch * mm);
    if (noise) d_noise.alloc((size_t)batch * std::max<long>(tew, 1) * m);
    double total_ms = 0.;
    for (long m0 = 0; m0 < N; m0 += batch) {
        const long nb = std::min(batch, N - m0);
        d_y.upload(ic + (size_t)m0 * n, (size_t)nb * n, st);
        d_q.upload(q0 + (size_t)m0 * nm, (size_t)nb * nm, st);
        d_am.upload(am0 + (size_t)m0 * mm, (size_t)nb * mm, st);
        if (r0) d_r0.upload(r0 + (size_t)m0 * mm, (size_t)nb * mm, st);
        if (noise && tew) d_noise.upload(noise + (size_t)m0 * tew * m, (size_t)nb * tew * m, st);
        // parts 1-3: forward pass, trajectory following the micro steps (mode 2), every step recorded
        TgParams P;
        benettin_fill_common(P, tab, nb, m, 0, 1.);
        P.forward = 2;
        P.n_pre = n_pre;
        P.n_rec = tew;
        P.dt_macro = d_dtm.p;
        P.sub_ptr = d_ptr.p;
        P.sub_dt = d_sub.p;
        P.write_steps = 1;
        P.n_records = Rf;
        P.y = d_y.p;
        P.fm = d_q.p;
        P.rec_y = d_ry.p;
        P.rec_fm = d_rq.p;
        P.rec_exp = d_re.p;
        P.r0 = r0 ? d_r0.p : nullptr;
        P.r_all = d_rall.p;
        P.r_first = n_pre;
        QGSB_CUDA(cudaEventRecord(cx.ev0, st));
        benettin_dispatch(t, tab, P, scratch);
        // parts 4-5: backward recursion
        GinelliParams G;
        G.n_members = nb;
        G.n = n;
        G.m = m;
        G.tw = tw;
        G.tew = tew;
        G.write_steps = write_steps;
        G.n_records = R;
        G.noise_pert = noise_pert;
        G.r_all = d_rall.p;
        G.rec_y = d_ry.p;
        G.rec_q = d_rq.p;
        G.am0 = d_am.p;
        G.noise = noise ? d_noise.p : nullptr;
        G.dte = d_dte.p;
        G.out_y = d_oy.p;
        G.out_exp = d_oe.p;
        G.out_vec = d_ov.p;
        const size_t bytes = (mm + (size_t)m * (m + 1)) * sizeof(double);
        QGSB_REQUIRE(bytes <= cx.smem_optin, "n_vec = %d: the backward recursion keeps two n_vec x n_vec matrices in "
                     "shared memory (%zu bytes > %zu)", m, bytes, cx.smem_optin);
        if (bytes > 48 * 1024)
            QGSB_CUDA(cudaFuncSetAttribute(ginelli_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        ginelli_kernel<<<(unsigned)nb, std::max(32, (m + 31) / 32 * 32), bytes, st>>>(G);   // thread = column of A
        count_launch();
        QGSB_CUDA(cudaGetLastError());
        QGSB_CUDA(cudaEventRecord(cx.ev1, st));
        d_oy.download(rec_traj + (size_t)m0 * n * R, (size_t)nb * n * R, st);
        d_oe.download(rec_exp + (size_t)m0 * m * R, (size_t)nb * m * R, st);
        d_ov.download(rec_vec + (size_t)m0 * nm * R, (size_t)nb * nm * R, st);
        QGSB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        total_ms += ms;
    }
    if (device_ms) *device_ms = total_ms;
}

extern "C" int qgsb_clv_ginelli(const qgsb_tensor *t, long N, const double *ic, int n_vec, const double *q0,
                                const double *r0, long n_pre, long n_time, long n_after, const double *dt_macro,
                                const long *sub_ptr, const double *sub_dt, int s, const double *a, const double *b,
                                const double *c, long write_steps, double noise_pert, const double *am0,
                                const double *noise, const double *dte, long R, double *rec_traj, double *rec_exp,
                                double *rec_vec, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && ic && q0 && dt_macro && sub_ptr && sub_dt && am0 && dte && rec_traj && rec_exp && rec_vec,
                 "null argument");
    QGSB_REQUIRE(N >= 1 && n_pre >= 0 && n_time >= 0 && n_after >= 0 && write_steps >= 0, "bad sizes");
    QGSB_REQUIRE(n_vec >= 1 && n_vec <= t->view.n && n_vec <= 1024, "n_vec must be in 1..min(n_dim, 1024)");
    QGSB_REQUIRE(t->jnnz_in > 0, "tensor handle has no Jacobian tensor");
    QGSB_REQUIRE(noise_pert == 0. || noise != nullptr, "noise_pert needs a noise array");
    ensure_init();
    // members are independent (every one carries its own basis and recursion): contiguous blocks, one per device
    const int n = t->view.n, m = n_vec;
    const size_t nm = (size_t)n * m, mm = (size_t)m * m;
    const long tew = n_time + n_after;
    const int parts = shard_count(N, 512);
    std::vector<double> ms(parts, 0.);
    run_sharded(N, parts, [&](int g, long lo, long hi) {
        ginelli_device(tensor_here(t), hi - lo, ic + (size_t)lo * n, n_vec, q0 + (size_t)lo * nm,
                       r0 ? r0 + (size_t)lo * mm : nullptr, n_pre, n_time, n_after, dt_macro, sub_ptr, sub_dt, s, a, b,
                       write_steps, noise_pert, am0 + (size_t)lo * mm, noise ? noise + (size_t)lo * tew * m : nullptr,
                       dte, R, rec_traj + (size_t)lo * n * R, rec_exp + (size_t)lo * m * R,
                       rec_vec + (size_t)lo * nm * R, &ms[g]);
    });
    if (device_ms) *device_ms = *std::max_element(ms.begin(), ms.end());
    QGSB_API_END
}

extern "C" int qgsb_clv_subspace_intersect(long N, int n, long R, const double *bvec, const double *fvec, double *vec)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(bvec && fvec && vec, "null argument");
    QGSB_REQUIRE(N >= 1 && R >= 1 && n >= 1 && n <= 63, "sizes out of range (n_dim <= 63)");
    ensure_init();
    Context &cx = ctx();
    cudaStream_t st = cx.stream;
    const size_t per_member = (size_t)n * n * R;
    const size_t budget = std::min<size_t>(cx.total_mem / 4, (size_t)24 << 30);
    const long batch = std::max<long>(1, std::min<long>(N, (long)(budget / (3 * per_member * sizeof(double)))));
    DevBuf<double> d_b((size_t)batch * per_member), d_f((size_t)batch * per_member), d_v((size_t)batch * per_member);
    const size_t bytes = (4 * (size_t)n * (n + 1) + (n + 2)) * sizeof(double);
    if (bytes > 48 * 1024)
        QGSB_CUDA(cudaFuncSetAttribute(subspace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    for (long m0 = 0; m0 < N; m0 += batch) {
        const long nb = std::min(batch, N - m0);
        d_b.upload(bvec + (size_t)m0 * per_member, (size_t)nb * per_member, st);
        d_f.upload(fvec + (size_t)m0 * per_member, (size_t)nb * per_member, st);
        SubspaceParams P;
        P.n_pairs = nb * R;
        P.R = R;
        P.n = n;
        P.bvec = d_b.p;
        P.fvec = d_f.p;
        P.vec = d_v.p;
        subspace_kernel<<<(unsigned)P.n_pairs, 64, bytes, st>>>(P);
        count_launch();
        QGSB_CUDA(cudaGetLastError());
        d_v.download(vec + (size_t)m0 * per_member, (size_t)nb * per_member, st);
        QGSB_CUDA(cudaStreamSynchronize(st));
    }
    QGSB_API_END
}
