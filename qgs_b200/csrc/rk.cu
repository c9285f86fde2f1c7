// rk.cu -- fused explicit Runge-Kutta ensemble kernels (generic tensors) and their C entry points.
//
// Replaces _integrate_runge_kutta_jit (qgs/integrators/integrate.py:182-223) and the
// TrajectoryProcess pool that fans it over members (qgs/integrators/integrator.py:453-512).
// One launch integrates every member over all time steps and all stages; the tensor is staged once
// into shared memory and reused for every evaluation of f.
#include <algorithm>
#include <cmath>
#include <numeric>

#include "common.cuh"
#include "kernels.cuh"
#include "spec_registry.h"
#include "tgls_shared.cuh"

namespace qgsb {

// ------------------------------------------------------------------------------------------------
Tableau make_tableau(int s, const double *a, const double *b)
{
    QGSB_REQUIRE(s >= 1 && s <= 16, "number of Runge-Kutta stages must be in 1..16, got %d", s);
    QGSB_REQUIRE(a && b, "null Butcher tableau");
    Tableau t;
    t.s = s;
    t.a.assign((size_t)s * s, 0.);
    t.b.assign(b, b + s);
    t.chain = true;
    t.alpha.assign(s, 0.);
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < i; ++j) {  // k[j >= i] is still zero when stage i is formed (integrate.py:214-217)
            t.a[(size_t)i * s + j] = a[(size_t)i * s + j];
            if (a[(size_t)i * s + j] != 0.) {
                if (j == i - 1) t.alpha[i] = a[(size_t)i * s + j];
                else t.chain = false;
            }
        }
    return t;
}

struct RkParams {
    long ld;          // member stride of the state / record arrays
    long n_members;
    long n_steps;
    const double *dt; // (n_steps) device
    long write_steps;
    long n_records;
    double *rec;      // (R, n, ld) or nullptr
    double *y;        // (n, ld) in/out
    int s;
    int chain;
    int tensor_in_smem;
    double a[16 * 16];
    double b[16];
    double alpha[16];
};

// ------------------------------------------------------------------------------------------------
// G1: one thread per member, state in shared memory as [variable][thread] (conflict-free), tensor
// entries broadcast to the warp from shared memory (or L1 when the tensor is too large).
// ------------------------------------------------------------------------------------------------
template <int RANK, int BS>
__device__ __forceinline__ double row_eval(const Entry *__restrict__ ent, int e0, int e1,
                                           const double *__restrict__ xs)
{
    double acc = 0.;
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Entry en = ent[e];
        double p = xs[(en.jk & 0xffffu) * BS] * xs[(en.jk >> 16) * BS];
        if (RANK == 5) p = p * xs[(en.lm & 0xffffu) * BS] * xs[(en.lm >> 16) * BS];
        acc += p * en.v;  // (a*b)*value, then +=   (sparse_mul.py:79)
    }
    return acc;
}

template <int RANK, int BS>
__global__ void __launch_bounds__(BS) rk_g1_kernel(TensorView T, const __grid_constant__ RkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = T.n, tid = threadIdx.x, s = P.s;
    const long member = (long)blockIdx.x * BS + tid;

    // ---- carve shared memory ----
    Entry *s_ent = reinterpret_cast<Entry *>(smem_raw);
    size_t off = P.tensor_in_smem ? sizeof(Entry) * (size_t)T.nnz : 0;
    double *xs = reinterpret_cast<double *>(smem_raw + off);        // (n+1, BS)
    double *y = xs + (size_t)(n + 1) * BS;                           // (n, BS)
    double *buf2 = y + (size_t)n * BS;                               // chain: xn (n+1, BS); general: K (s, n, BS)
    double *acc = buf2 + (size_t)(n + 1) * BS;                       // chain only: (n, BS)
    int *s_row = reinterpret_cast<int *>(buf2 + (P.chain ? (size_t)(2 * n + 1) * BS : (size_t)s * n * BS));

    const Entry *ent = T.ent;
    if (P.tensor_in_smem) {
        const int4 *src = reinterpret_cast<const int4 *>(T.ent);
        int4 *dst = reinterpret_cast<int4 *>(s_ent);
        for (int e = tid; e < T.nnz; e += BS) dst[e] = src[e];
        ent = s_ent;
    }
    for (int i = tid; i < n + 2; i += BS) s_row[i] = T.row_ptr[i];
    xs += tid;
    y += tid;
    buf2 += tid;
    acc += tid;
    const size_t gbase = tile_base(member, n);
    for (int i = 0; i < n; ++i) {
        double v = P.y[gbase + (size_t)i * TILE];
        y[(size_t)i * BS] = v;
        xs[(size_t)(i + 1) * BS] = v;
    }
    xs[0] = 1.;
    if (P.chain) buf2[0] = 1.;
    __syncthreads();

    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {       // integrate.py:210-212
            double *r = P.rec + (size_t)iw * n * P.ld + gbase;
            for (int i = 0; i < n; ++i) r[(size_t)i * TILE] = y[(size_t)i * BS];
            ++iw;
        }
        if (P.chain) {
            double *xin = xs, *xout = buf2;
            for (int st = 0; st < s; ++st) {
                const double wb = dt * P.b[st];
                const double wa = st + 1 < s ? dt * P.alpha[st + 1] : 0.;
                const bool last = st + 1 == s;
                for (int i = 1; i <= n; ++i) {
                    const double k = row_eval<RANK, BS>(ent, s_row[i], s_row[i + 1], xin);
                    const size_t o = (size_t)(i - 1) * BS;
                    const double ac = st == 0 ? wb * k : acc[o] + wb * k;
                    if (!last) {
                        acc[o] = ac;
                        xout[(size_t)i * BS] = y[o] + wa * k;              // y + (dt a[i]) @ k   integrate.py:216
                    } else {
                        const double yn = y[o] + ac;                        // y + (dt b) @ k      integrate.py:218
                        y[o] = yn;
                        xout[(size_t)i * BS] = yn;
                    }
                }
                double *tmp = xin;
                xin = xout;
                xout = tmp;
            }
            if (s & 1) {  // keep the convention "xs holds the current state" for the next step
                for (int i = 1; i <= n; ++i) xs[(size_t)i * BS] = y[(size_t)(i - 1) * BS];
            }
        } else {
            double *K = buf2;
            for (int st = 0; st < s; ++st) {
                for (int i = 1; i <= n; ++i) {
                    double v = 0.;
                    for (int j = 0; j < st; ++j) {
                        const double w = P.a[st * s + j];
                        if (w != 0.) v += (dt * w) * K[((size_t)j * n + (i - 1)) * BS];
                    }
                    xs[(size_t)i * BS] = y[(size_t)(i - 1) * BS] + v;
                }
                for (int i = 1; i <= n; ++i)
                    K[((size_t)st * n + (i - 1)) * BS] = row_eval<RANK, BS>(ent, s_row[i], s_row[i + 1], xs);
            }
            for (int i = 0; i < n; ++i) {
                double v = 0.;
                for (int j = 0; j < s; ++j) v += (dt * P.b[j]) * K[((size_t)j * n + i) * BS];
                y[(size_t)i * BS] += v;
            }
        }
    }
    if (P.rec) {                                                            // integrate.py:221
        double *r = P.rec + (size_t)(P.n_records - 1) * n * P.ld + gbase;
        for (int i = 0; i < n; ++i) r[(size_t)i * TILE] = y[(size_t)i * BS];
    }
    for (int i = 0; i < n; ++i) P.y[gbase + (size_t)i * TILE] = y[(size_t)i * BS];
}

// ------------------------------------------------------------------------------------------------
// G2: one warp per member for large bases (state does not fit a thread's share of shared memory).
// Lanes stride over the entries of a row (coalesced 512 B reads of the tensor stream) and the row
// sum is reduced by warp shuffles.
// ------------------------------------------------------------------------------------------------
template <int RANK>
__device__ __forceinline__ double row_eval_warp(const Entry *__restrict__ ent, int e0, int e1,
                                                const double *__restrict__ xs, int lane)
{
    double acc = 0.;
    for (int e = e0 + lane; e < e1; e += 32) {
        const Entry en = ent[e];
        double p = xs[en.jk & 0xffffu] * xs[en.jk >> 16];
        if (RANK == 5) p = p * xs[en.lm & 0xffffu] * xs[en.lm >> 16];
        acc += p * en.v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

constexpr int G2_WARPS = 4;

template <int RANK>
__global__ void __launch_bounds__(G2_WARPS * 32) rk_g2_kernel(TensorView T, const __grid_constant__ RkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = T.n, s = P.s, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long member = (long)blockIdx.x * G2_WARPS + warp;
    const size_t per_member = P.chain ? (size_t)(4 * n + 2) : (size_t)(2 * n + 1) + (size_t)s * n;
    double *xs = reinterpret_cast<double *>(smem_raw) + per_member * warp;  // (n+1)
    double *y = xs + (n + 1);                                               // (n)
    double *buf2 = y + n;                                                    // chain: xn (n+1); general: K (s, n)
    double *acc = buf2 + (n + 1);                                            // chain only (n)
    const bool active = member < P.ld;
    const int *row = T.row_ptr;
    const Entry *ent = T.ent;

    const size_t gbase = tile_base(member, n);
    if (active) {
        for (int i = lane; i < n; i += 32) {
            double v = P.y[gbase + (size_t)i * TILE];
            y[i] = v;
            xs[i + 1] = v;
        }
        if (lane == 0) {
            xs[0] = 1.;
            if (P.chain) buf2[0] = 1.;
        }
    }
    __syncwarp();
    if (!active) return;

    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {
            double *r = P.rec + (size_t)iw * n * P.ld + gbase;
            for (int i = lane; i < n; i += 32) r[(size_t)i * TILE] = y[i];
            ++iw;
        }
        if (P.chain) {
            double *xin = xs, *xout = buf2;
            for (int st = 0; st < s; ++st) {
                const double wb = dt * P.b[st];
                const double wa = st + 1 < s ? dt * P.alpha[st + 1] : 0.;
                const bool last = st + 1 == s;
                for (int i = 1; i <= n; ++i) {
                    const double k = row_eval_warp<RANK>(ent, row[i], row[i + 1], xin, lane);
                    if (lane == 0) {
                        const double ac = st == 0 ? wb * k : acc[i - 1] + wb * k;
                        if (!last) {
                            acc[i - 1] = ac;
                            xout[i] = y[i - 1] + wa * k;
                        } else {
                            const double yn = y[i - 1] + ac;
                            y[i - 1] = yn;
                            xout[i] = yn;
                        }
                    }
                }
                __syncwarp();
                double *tmp = xin;
                xin = xout;
                xout = tmp;
            }
            if (s & 1) {
                for (int i = lane; i < n; i += 32) xs[i + 1] = y[i];
                __syncwarp();
            }
        } else {
            double *K = buf2;
            for (int st = 0; st < s; ++st) {
                for (int i = lane; i < n; i += 32) {
                    double v = 0.;
                    for (int j = 0; j < st; ++j) {
                        const double w = P.a[st * s + j];
                        if (w != 0.) v += (dt * w) * K[(size_t)j * n + i];
                    }
                    xs[i + 1] = y[i] + v;
                }
                __syncwarp();
                for (int i = 1; i <= n; ++i) {
                    const double k = row_eval_warp<RANK>(ent, row[i], row[i + 1], xs, lane);
                    if (lane == 0) K[(size_t)st * n + (i - 1)] = k;
                }
                __syncwarp();
            }
            for (int i = lane; i < n; i += 32) {
                double v = 0.;
                for (int j = 0; j < s; ++j) v += (dt * P.b[j]) * K[(size_t)j * n + i];
                y[i] += v;
            }
            __syncwarp();
        }
    }
    if (P.rec) {
        double *r = P.rec + (size_t)(P.n_records - 1) * n * P.ld + gbase;
        for (int i = lane; i < n; i += 32) r[(size_t)i * TILE] = y[i];
    }
    for (int i = lane; i < n; i += 32) P.y[gbase + (size_t)i * TILE] = y[i];
}

// ------------------------------------------------------------------------------------------------
// G3: large bases (n > 64, rank 3, chain tableaux).  A block owns MB members (a multiple of 32) and
// keeps only their stage state x in shared memory ([variable][member], conflict-free, ~176 KB => 96
// members per SM for the 228-variable 6x6 model); y, the weighted stage sum and the next stage state
// stay in the tiled-SoA arrays in global memory, where each value is touched once per stage with
// coalesced accesses (40 n bytes per member and stage, L2 resident; the loads of a row are issued
// before the row's terms so that their latency is covered).
// To have enough warps in flight the rows are split into G groups of about equal entry counts and
// thread (h, m) evaluates the rows of group h for member m: G x MB threads, all reading the same x.
// The tensor is a 16-byte-per-entry stream walked identically by the MB threads of a group: the group
// pulls its part through a double buffer of chunks (<= 256 entries, whole rows padded to multiples
// of four entries) with cp.async, one chunk ahead, one named barrier per chunk, and consumes it with
// broadcast LDS.128 in a software pipeline, four independent partial sums per row.  An entry holds
// its value and the byte offsets of x_j / x_k in the shared state (premultiplied by MB).
// The kernel is bound by shared-memory bandwidth: 16 bytes of x + 16/32 bytes of entry per term and member.
// ------------------------------------------------------------------------------------------------
struct __align__(16) G3Entry {
    double v;
    uint32_t oj, ok;   // byte offsets of the x_j / x_k rows in the shared state
};

struct G3Chunk {
    int off, count, row0, nrows;   // entries [off, off + count) = rows row0 .. row0 + nrows - 1 (0-based)
};

constexpr int G3_CHUNK = 256;               // entries per chunk
constexpr int G3_STRIDE = G3_CHUNK + 4;     // the software pipeline reads one group of four past the end of a chunk
constexpr int G3_RING = 2;
constexpr int G3_MAX_GROUPS = 8;

struct G3Plan {
    int groups;
    int chunk0[G3_MAX_GROUPS + 1];   // group h owns chunks [chunk0[h], chunk0[h + 1])
    int row0[G3_MAX_GROUPS + 1];     // and rows [row0[h], row0[h + 1])
};

__device__ __forceinline__ void g3_fetch(G3Entry *ring, const G3Entry *__restrict__ ent, const G3Chunk *chunks,
                                         long chunk, int per_stage, int m, int mb)
{
    // the chunk counter runs on across stages and steps; every stage walks the same list
    const G3Chunk ch = chunks[chunk % per_stage];
    G3Entry *dst = ring + (size_t)(chunk % G3_RING) * G3_STRIDE;
    const G3Entry *src = ent + ch.off;
    for (int e = m; e < ch.count; e += mb) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + e);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + e));
    }
    asm volatile("cp.async.commit_group;");
}

#define G3_X(off) (*reinterpret_cast<const double *>(xb + (off)))

__global__ void __launch_bounds__(512)
rk_g3_kernel(int n, int mb, int n_chunks, const __grid_constant__ G3Plan plan, const G3Entry *__restrict__ ent,
             const G3Chunk *__restrict__ chunks_g, const int *__restrict__ rowlen_g,
             const __grid_constant__ RkParams P, double *__restrict__ acc_g, double *__restrict__ xn_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, s = P.s;
    const int h = tid / mb, m = tid - h * mb;          // row group, member inside the block
    G3Entry *rings = reinterpret_cast<G3Entry *>(smem_raw);
    G3Chunk *chunks_all = reinterpret_cast<G3Chunk *>(rings + (size_t)G3_STRIDE * G3_RING * plan.groups);
    int *rowlen = reinterpret_cast<int *>(chunks_all + n_chunks);
    unsigned char *xb = reinterpret_cast<unsigned char *>(rowlen + ((n + 3) & ~3)) + (size_t)m * 8;  // x_i at xb + i mb 8
    const size_t rowb = (size_t)mb * 8;
    long member = (long)blockIdx.x * mb + m;
    const bool act = member < P.ld;                   // the whole block keeps running: it shares rings and barriers
    if (!act) member = P.ld - 1;
    const size_t gbase = tile_base(member, n);
    double *yg = P.y + gbase, *ag = acc_g + gbase, *xg = xn_g + gbase;
    G3Entry *ring = rings + (size_t)G3_STRIDE * G3_RING * h;
    const G3Chunk *chunks = chunks_all + plan.chunk0[h];
    const int per_stage = plan.chunk0[h + 1] - plan.chunk0[h];
    const int r_lo = plan.row0[h], r_hi = plan.row0[h + 1];
    for (int q = tid; q < G3_STRIDE * G3_RING * plan.groups; q += blockDim.x)
        rings[q] = G3Entry{0., 0u, 0u};                // read-ahead lands on valid offsets
    for (int q = tid; q < n_chunks; q += blockDim.x) chunks_all[q] = chunks_g[q];
    for (int q = tid; q < n; q += blockDim.x) rowlen[q] = rowlen_g[q];
    if (h == 0) *reinterpret_cast<double *>(xb) = 1.;
    for (int i = r_lo; i < r_hi; ++i) *reinterpret_cast<double *>(xb + (size_t)(i + 1) * rowb) = yg[(size_t)i * TILE];
    __syncthreads();

    const long total = (long)per_stage * s * P.n_steps;
    long chunk = 0;
    if (total > 0) g3_fetch(ring, ent, chunks, 0, per_stage, m, mb);

    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {       // integrate.py:210-212
            if (act) {
                double *r = P.rec + (size_t)iw * n * P.ld + gbase;
                for (int i = r_lo; i < r_hi; ++i) r[(size_t)i * TILE] = yg[(size_t)i * TILE];
            }
            ++iw;
        }
        for (int st = 0; st < s; ++st) {
            const double wb = dt * P.b[st];
            const double wa = st + 1 < s ? dt * P.alpha[st + 1] : 0.;
            const bool last = st + 1 == s;
            for (int c = 0; c < per_stage; ++c, ++chunk) {
                asm volatile("cp.async.wait_group 0;");
                asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"(mb));   // chunk landed for the group; chunk - 1 consumed
                if (chunk + 1 < total) g3_fetch(ring, ent, chunks, chunk + 1, per_stage, m, mb);
                const uint4 *pe = reinterpret_cast<const uint4 *>(ring + (size_t)(chunk % G3_RING) * G3_STRIDE);
                const G3Chunk ch = chunks[c];
                // software pipeline over groups of four entries: the entries of the next group are loaded while the
                // x values of the current one are in flight (rows are padded to whole groups with zero entries)
                uint4 c0 = pe[0], c1 = pe[1], c2 = pe[2], c3 = pe[3];
                for (int r = 0; r < ch.nrows; ++r) {
                    const int groups = rowlen[ch.row0 + r] >> 2;
                    const size_t o = (size_t)(ch.row0 + r) * TILE;
                    const double y_old = yg[o];                             // in flight while the row is summed
                    const double a_old = st == 0 ? 0. : ag[o];
                    double k0 = 0., k1 = 0., k2 = 0., k3 = 0.;
#pragma unroll 2
                    for (int g = 0; g < groups; ++g) {
                        const double a0 = G3_X(c0.z), b0 = G3_X(c0.w), a1 = G3_X(c1.z), b1 = G3_X(c1.w);
                        const double a2 = G3_X(c2.z), b2 = G3_X(c2.w), a3 = G3_X(c3.z), b3 = G3_X(c3.w);
                        pe += 4;
                        const uint4 n0 = pe[0], n1 = pe[1], n2 = pe[2], n3 = pe[3];
                        k0 = fma(a0 * b0, __hiloint2double(c0.y, c0.x), k0);   // (a*b)*value, then +=  (sparse_mul.py:79)
                        k1 = fma(a1 * b1, __hiloint2double(c1.y, c1.x), k1);
                        k2 = fma(a2 * b2, __hiloint2double(c2.y, c2.x), k2);
                        k3 = fma(a3 * b3, __hiloint2double(c3.y, c3.x), k3);
                        c0 = n0;
                        c1 = n1;
                        c2 = n2;
                        c3 = n3;
                    }
                    const double k = (k0 + k1) + (k2 + k3);
                    if (act) {
                        const double ac = a_old + wb * k;
                        if (!last) {
                            ag[o] = ac;
                            xg[o] = y_old + wa * k;                         // y + (dt a[i]) @ k   integrate.py:216
                        } else {
                            const double yn = y_old + ac;                   // y + (dt b) @ k      integrate.py:218
                            yg[o] = yn;
                            xg[o] = yn;
                        }
                    }
                }
            }
            __syncthreads();               // every group is done reading x
            for (int i = r_lo; i < r_hi; ++i)
                *reinterpret_cast<double *>(xb + (size_t)(i + 1) * rowb) = xg[(size_t)i * TILE];
            __syncthreads();               // the next stage state is complete
        }
    }
    asm volatile("cp.async.wait_group 0;");
    if (P.rec && act) {                                                     // integrate.py:221
        double *r = P.rec + (size_t)(P.n_records - 1) * n * P.ld + gbase;
        for (int i = r_lo; i < r_hi; ++i) r[(size_t)i * TILE] = yg[(size_t)i * TILE];
    }
}
#undef G3_X

// G3, two members per thread: a broadcast read of a 16-byte entry costs a warp as many shared-memory wavefronts as
// its two 8-byte reads of x (one wavefront per 32-bit word and warp, broadcast or not), so half of G3's traffic is
// the entry stream.  Here thread mt of a row group owns members 2 mt and 2 mt + 1 -- their x_j sit side by side and
// come in one LDS.128 -- and an entry is read once for 64 members instead of 32: 6 instead of 8 wavefronts per
// term and 32 members.  Twice the rows groups (8) keep as many warps in flight with 64 members per block.
#define G3_X2(off) (*reinterpret_cast<const double2 *>(xb + (off)))
__global__ void __launch_bounds__(512)
rk_g3p_kernel(int n, int mb, int n_chunks, const __grid_constant__ G3Plan plan, const G3Entry *__restrict__ ent,
              const G3Chunk *__restrict__ chunks_g, const int *__restrict__ rowlen_g,
              const __grid_constant__ RkParams P, double *__restrict__ acc_g, double *__restrict__ xn_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, s = P.s, tg = mb >> 1;
    const int h = tid / tg, mt = tid - h * tg;        // row group, member pair inside the block
    G3Entry *rings = reinterpret_cast<G3Entry *>(smem_raw);
    G3Chunk *chunks_all = reinterpret_cast<G3Chunk *>(rings + (size_t)G3_STRIDE * G3_RING * plan.groups);
    int *rowlen = reinterpret_cast<int *>(chunks_all + n_chunks);
    unsigned char *xb = reinterpret_cast<unsigned char *>(rowlen + ((n + 3) & ~3)) + (size_t)mt * 16;  // pair of x_i at xb + i mb 8
    const size_t rowb = (size_t)mb * 8;
    long member = (long)blockIdx.x * mb + 2 * mt;     // even: the pair never straddles a tile, and is 16-byte aligned
    const bool act = member < P.ld;
    if (!act) member = P.ld - 2;
    const size_t gbase = tile_base(member, n);
    double *yg = P.y + gbase, *ag = acc_g + gbase, *xg = xn_g + gbase;
    G3Entry *ring = rings + (size_t)G3_STRIDE * G3_RING * h;
    const G3Chunk *chunks = chunks_all + plan.chunk0[h];
    const int per_stage = plan.chunk0[h + 1] - plan.chunk0[h];
    const int r_lo = plan.row0[h], r_hi = plan.row0[h + 1];
    for (int q = tid; q < G3_STRIDE * G3_RING * plan.groups; q += blockDim.x)
        rings[q] = G3Entry{0., 0u, 0u};                // read-ahead lands on valid offsets
    for (int q = tid; q < n_chunks; q += blockDim.x) chunks_all[q] = chunks_g[q];
    for (int q = tid; q < n; q += blockDim.x) rowlen[q] = rowlen_g[q];
    if (h == 0) *reinterpret_cast<double2 *>(xb) = make_double2(1., 1.);
    for (int i = r_lo; i < r_hi; ++i)
        *reinterpret_cast<double2 *>(xb + (size_t)(i + 1) * rowb) = *reinterpret_cast<const double2 *>(yg + (size_t)i * TILE);
    __syncthreads();

    const long total = (long)per_stage * s * P.n_steps;
    long chunk = 0;
    if (total > 0) g3_fetch(ring, ent, chunks, 0, per_stage, mt, tg);

    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {       // integrate.py:210-212
            if (act) {
                double *r = P.rec + (size_t)iw * n * P.ld + gbase;
                for (int i = r_lo; i < r_hi; ++i)
                    *reinterpret_cast<double2 *>(r + (size_t)i * TILE) = *reinterpret_cast<const double2 *>(yg + (size_t)i * TILE);
            }
            ++iw;
        }
        for (int st = 0; st < s; ++st) {
            const double wb = dt * P.b[st];
            const double wa = st + 1 < s ? dt * P.alpha[st + 1] : 0.;
            const bool last = st + 1 == s;
            for (int c = 0; c < per_stage; ++c, ++chunk) {
                asm volatile("cp.async.wait_group 0;");
                asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"(tg));   // chunk landed for the group; chunk - 1 consumed
                if (chunk + 1 < total) g3_fetch(ring, ent, chunks, chunk + 1, per_stage, mt, tg);
                const uint4 *pe = reinterpret_cast<const uint4 *>(ring + (size_t)(chunk % G3_RING) * G3_STRIDE);
                const G3Chunk ch = chunks[c];
                uint4 c0 = pe[0], c1 = pe[1], c2 = pe[2], c3 = pe[3];
                for (int r = 0; r < ch.nrows; ++r) {
                    const int groups = rowlen[ch.row0 + r] >> 2;
                    const size_t o = (size_t)(ch.row0 + r) * TILE;
                    const double2 y_old = *reinterpret_cast<const double2 *>(yg + o);   // in flight while the row is summed
                    const double2 a_old = st == 0 ? make_double2(0., 0.) : *reinterpret_cast<const double2 *>(ag + o);
                    double k0 = 0., k1 = 0., k2 = 0., k3 = 0., l0 = 0., l1 = 0., l2 = 0., l3 = 0.;
                    double2 a3 = make_double2(0., 0.);
#pragma unroll 2
                    for (int g = 0; g < groups; ++g) {
                        // bit 0 of the first offset: same x_j as the entry before (entries are sorted by (i, j, k)), so
                        // the value is carried in registers instead of read again -- a warp-uniform predicate
                        double2 a0 = a3, a1, a2;
                        if (!(c0.z & 1u)) a0 = G3_X2(c0.z);
                        a1 = a0;
                        if (!(c1.z & 1u)) a1 = G3_X2(c1.z);
                        a2 = a1;
                        if (!(c2.z & 1u)) a2 = G3_X2(c2.z);
                        a3 = a2;
                        if (!(c3.z & 1u)) a3 = G3_X2(c3.z);
                        const double2 b0 = G3_X2(c0.w), b1 = G3_X2(c1.w), b2 = G3_X2(c2.w), b3 = G3_X2(c3.w);
                        pe += 4;
                        const uint4 n0 = pe[0], n1 = pe[1], n2 = pe[2], n3 = pe[3];
                        const double v0 = __hiloint2double(c0.y, c0.x), v1 = __hiloint2double(c1.y, c1.x);
                        const double v2 = __hiloint2double(c2.y, c2.x), v3 = __hiloint2double(c3.y, c3.x);
                        k0 = fma(a0.x * b0.x, v0, k0);                         // (a*b)*value, then +=  (sparse_mul.py:79)
                        l0 = fma(a0.y * b0.y, v0, l0);
                        k1 = fma(a1.x * b1.x, v1, k1);
                        l1 = fma(a1.y * b1.y, v1, l1);
                        k2 = fma(a2.x * b2.x, v2, k2);
                        l2 = fma(a2.y * b2.y, v2, l2);
                        k3 = fma(a3.x * b3.x, v3, k3);
                        l3 = fma(a3.y * b3.y, v3, l3);
                        c0 = n0;
                        c1 = n1;
                        c2 = n2;
                        c3 = n3;
                    }
                    const double k = (k0 + k1) + (k2 + k3), l = (l0 + l1) + (l2 + l3);
                    if (act) {
                        const double2 ac = make_double2(a_old.x + wb * k, a_old.y + wb * l);
                        if (!last) {
                            *reinterpret_cast<double2 *>(ag + o) = ac;
                            *reinterpret_cast<double2 *>(xg + o) = make_double2(y_old.x + wa * k, y_old.y + wa * l);
                        } else {
                            const double2 yn = make_double2(y_old.x + ac.x, y_old.y + ac.y);
                            *reinterpret_cast<double2 *>(yg + o) = yn;
                            *reinterpret_cast<double2 *>(xg + o) = yn;
                        }
                    }
                }
            }
            __syncthreads();               // every group is done reading x
            for (int i = r_lo; i < r_hi; ++i)
                *reinterpret_cast<double2 *>(xb + (size_t)(i + 1) * rowb) = *reinterpret_cast<const double2 *>(xg + (size_t)i * TILE);
            __syncthreads();               // the next stage state is complete
        }
    }
    asm volatile("cp.async.wait_group 0;");
    if (P.rec && act) {                                                     // integrate.py:221
        double *r = P.rec + (size_t)(P.n_records - 1) * n * P.ld + gbase;
        for (int i = r_lo; i < r_hi; ++i)
            *reinterpret_cast<double2 *>(r + (size_t)i * TILE) = *reinterpret_cast<const double2 *>(yg + (size_t)i * TILE);
    }
}
#undef G3_X2

// ------------------------------------------------------------------------------------------------
// Rows kernel: the latency regime (few members, many steps -- the reference's example scripts integrate ONE
// trajectory, qgs_maooam.py:100-117).  The throughput kernels give a member to one thread, which then walks all n
// rows one after the other (3 us per MAOOAM-36 step however few members there are).  Here a block owns one member
// and thread r owns ROW r: y_r and the weighted stage sum stay in its registers, the stage state is exchanged
// through a double buffer in shared memory (one barrier per stage), and the row is summed from the ELL table
// (staged in shared memory) with two independent partial sums.  Chain tableaux, ELL-able tensors.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rows_at(const double *base, unsigned off)
{
    return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + off);
}

// EFU = entries per row rounded up (template, so the row's entries live in registers for the whole integration and
// the sums are fully unrolled: per stage a thread issues its 2 EFU independent loads of x back to back)
template <int RANK, int EFU>
__global__ void __launch_bounds__(256) rk_rows_kernel(int n, const PEnt *__restrict__ f_ent, int EF,
                                                      const __grid_constant__ RkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = threadIdx.x, s = P.s;
    const long member = blockIdx.x;
    double *cur = reinterpret_cast<double *>(smem_raw);          // (n + 1), [0] = 1
    double *nxt = cur + ((n + 2) & ~1);
    const bool act = r < n;
    const size_t g = tile_base(member, n) + (size_t)(act ? r : 0) * TILE;
    double y = act ? P.y[g] : 0., acc = 0.;
    // this row's entries: value and the byte offsets of its factors
    double v[EFU];
    unsigned ab[EFU], cd[RANK == 5 ? EFU : 1];
#pragma unroll
    for (int q = 0; q < EFU; ++q) {
        PEnt e = {0., 0, 0, 0, 0};
        if (act && q < EF) e = f_ent[(size_t)q * n + r];
        v[q] = e.v;
        ab[q] = (unsigned)e.a | ((unsigned)e.b << 16);
        if (RANK == 5) cd[q] = (unsigned)e.c | ((unsigned)e.d << 16);
    }
    if (r == 0) {
        cur[0] = 1.;
        nxt[0] = 1.;
    }
    if (act) cur[r + 1] = y;
    __syncthreads();
    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {       // integrate.py:210-212
            if (act) P.rec[(size_t)iw * n * P.ld + g] = y;
            ++iw;
        }
        for (int st = 0; st < s; ++st) {
            double p[EFU];
#pragma unroll
            for (int q = 0; q < EFU; ++q) {
                p[q] = rows_at(cur, ab[q] & 0xffffu) * rows_at(cur, ab[q] >> 16);
                if (RANK == 5) p[q] = p[q] * rows_at(cur, cd[q] & 0xffffu) * rows_at(cur, cd[q] >> 16);
            }
            double k0 = 0., k1 = 0., k2 = 0., k3 = 0.;
#pragma unroll
            for (int q = 0; q < EFU; q += 4) {
                k0 = fma(p[q], v[q], k0);                                   // (a*b)*value, then +=  (sparse_mul.py:79)
                if (q + 1 < EFU) k1 = fma(p[q + 1], v[q + 1], k1);
                if (q + 2 < EFU) k2 = fma(p[q + 2], v[q + 2], k2);
                if (q + 3 < EFU) k3 = fma(p[q + 3], v[q + 3], k3);
            }
            const double k = (k0 + k1) + (k2 + k3);
            const double wb = dt * P.b[st];
            acc = st == 0 ? wb * k : acc + wb * k;
            if (st + 1 < s) {
                if (act) nxt[r + 1] = y + (dt * P.alpha[st + 1]) * k;       // y + (dt a[i]) @ k   integrate.py:216
            } else {
                y += acc;                                                   // y + (dt b) @ k      integrate.py:218
                if (act) nxt[r + 1] = y;
            }
            __syncthreads();
            double *t = cur;
            cur = nxt;
            nxt = t;
        }
    }
    if (act) {
        if (P.rec) P.rec[(size_t)(P.n_records - 1) * n * P.ld + g] = y;     // integrate.py:221
        P.y[g] = y;
    }
}

// ------------------------------------------------------------------------------------------------
// Wide kernel: the latency regime for large bases (a handful of trajectories of the 228-variable model).  One
// 1024-thread block owns one member; its 32 warps share the rows (assigned on the host by decreasing length so
// that every warp streams about the same number of entries), a warp's lanes stride over a row's entries
// (coalesced 512-byte reads of the tensor stream from L2) and reduce the row with a fixed shuffle tree, lane 0
// finishes the row.  The stage state is a double buffer in shared memory, one barrier per stage.
// 6x6, one member: 47 us per step, where the thread-per-member kernels need 1 ms and one CPU core of the reference
// 360 us; ncu: the random gathers of x_j, x_k from shared memory dominate (bank conflicts: short-scoreboard and MIO
// stalls, shared-memory wavefronts at 56 % of peak), then the tensor stream from L2.
// ------------------------------------------------------------------------------------------------
constexpr int WIDE_WARPS = 32;

template <int RANK>
__global__ void __launch_bounds__(WIDE_WARPS * 32) rk_wide_kernel(TensorView T, const int *__restrict__ warp_ptr,
                                                                   const int *__restrict__ warp_rows,
                                                                   const __grid_constant__ RkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = T.n, s = P.s, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long member = blockIdx.x;
    double *cur = reinterpret_cast<double *>(smem_raw);          // (n + 1), [0] = 1
    double *nxt = cur + ((n + 2) & ~1);
    double *ys = nxt + ((n + 2) & ~1);                           // (n)
    double *accs = ys + ((n + 1) & ~1);                          // (n)
    int *rinfo = reinterpret_cast<int *>(accs + ((n + 1) & ~1));  // (n, 3): row, first entry, end -- in warp order
    const size_t gbase = tile_base(member, n);
    for (int i = tid; i < n; i += blockDim.x) {
        const double v = P.y[gbase + (size_t)i * TILE];
        ys[i] = v;
        cur[i + 1] = v;
        const int row = warp_rows[i];
        rinfo[3 * i] = row;
        rinfo[3 * i + 1] = T.row_ptr[row];
        rinfo[3 * i + 2] = T.row_ptr[row + 1];
    }
    if (tid == 0) {
        cur[0] = 1.;
        nxt[0] = 1.;
    }
    __syncthreads();
    const int r0 = warp_ptr[warp], r1 = warp_ptr[warp + 1];
    long iw = 0;
    for (long ti = 0; ti < P.n_steps; ++ti) {
        const double dt = P.dt[ti];
        if (P.rec && P.write_steps > 0 && ti % P.write_steps == 0) {       // integrate.py:210-212
            double *rp = P.rec + (size_t)iw * n * P.ld + gbase;
            for (int i = tid; i < n; i += blockDim.x) rp[(size_t)i * TILE] = ys[i];
            ++iw;
        }
        for (int st = 0; st < s; ++st) {
            const double wb = dt * P.b[st];
            const double wa = st + 1 < s ? dt * P.alpha[st + 1] : 0.;
            const bool last = st + 1 == s;
            for (int q = r0; q < r1; ++q) {
                const int i = rinfo[3 * q], e1 = rinfo[3 * q + 2];          // 1-based row
                double k0 = 0., k1 = 0.;
                // 128 entries of the row are requested at once (4 independent 16-byte loads per lane)
                for (int e = rinfo[3 * q + 1] + lane; e < e1; e += 128) {
                    Entry en[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        en[u].v = 0.;
                        en[u].jk = 0u;
                        en[u].lm = 0u;
                        if (e + 32 * u < e1) en[u] = T.ent[e + 32 * u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double p = cur[en[u].jk & 0xffffu] * cur[en[u].jk >> 16];
                        if (RANK == 5) p = p * cur[en[u].lm & 0xffffu] * cur[en[u].lm >> 16];
                        if (u & 1) k1 = fma(p, en[u].v, k1);                // (a*b)*value, then +=  (sparse_mul.py:79)
                        else k0 = fma(p, en[u].v, k0);
                    }
                }
                double k = k0 + k1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
                if (lane == 0) {
                    const double ac = st == 0 ? wb * k : accs[i - 1] + wb * k;
                    if (!last) {
                        accs[i - 1] = ac;
                        nxt[i] = ys[i - 1] + wa * k;                        // y + (dt a[i]) @ k   integrate.py:216
                    } else {
                        const double yn = ys[i - 1] + ac;                   // y + (dt b) @ k      integrate.py:218
                        ys[i - 1] = yn;
                        nxt[i] = yn;
                    }
                }
            }
            __syncthreads();
            double *t = cur;
            cur = nxt;
            nxt = t;
        }
    }
    if (P.rec) {                                                            // integrate.py:221
        double *rp = P.rec + (size_t)(P.n_records - 1) * n * P.ld + gbase;
        for (int i = tid; i < n; i += blockDim.x) rp[(size_t)i * TILE] = ys[i];
    }
    for (int i = tid; i < n; i += blockDim.x) P.y[gbase + (size_t)i * TILE] = ys[i];
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
template <typename K>
static void set_smem(K kernel, size_t bytes)
{
    QGSB_REQUIRE(bytes <= ctx().smem_optin, "kernel needs %zu bytes of shared memory, device allows %zu", bytes,
                 ctx().smem_optin);
    if (bytes > 48 * 1024)
        QGSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

static void fill_params(RkParams &P, const Tableau &tab, double *d_y, long ld, long N, long n_steps,
                        const double *d_dt, long write_steps, long R, double *d_rec)
{
    memset(&P, 0, sizeof(P));
    P.ld = ld;
    P.n_members = N;
    P.n_steps = n_steps;
    P.dt = d_dt;
    P.write_steps = write_steps;
    P.n_records = R;
    P.rec = d_rec;
    P.y = d_y;
    P.s = tab.s;
    P.chain = tab.chain ? 1 : 0;
    for (int i = 0; i < tab.s; ++i) {
        P.b[i] = tab.b[i];
        P.alpha[i] = tab.alpha[i];
        for (int j = 0; j < tab.s; ++j) P.a[i * tab.s + j] = tab.a[(size_t)i * tab.s + j];
    }
}

template <int BS>
static void launch_g1(const qgsb_tensor *t, RkParams &P)
{
    const int n = t->view.n, s = P.s;
    const size_t per_thread = P.chain ? (size_t)(4 * n + 2) : (size_t)(2 * n + 1) + (size_t)s * n;
    const size_t ent_bytes = sizeof(Entry) * (size_t)t->view.nnz;
    size_t state = per_thread * 8 * BS + sizeof(int) * (n + 2);
    P.tensor_in_smem = (ent_bytes <= 32 * 1024 && state + ent_bytes <= ctx().smem_optin) ? 1 : 0;
    const size_t bytes = state + (P.tensor_in_smem ? ent_bytes : 0);
    const unsigned grid = (unsigned)(P.ld / BS);
    if (t->view.rank == 5) {
        set_smem(rk_g1_kernel<5, BS>, bytes);
        rk_g1_kernel<5, BS><<<grid, BS, bytes, ctx().stream>>>(t->view, P);
    } else {
        set_smem(rk_g1_kernel<3, BS>, bytes);
        rk_g1_kernel<3, BS><<<grid, BS, bytes, ctx().stream>>>(t->view, P);
    }
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

static void launch_g2(const qgsb_tensor *t, RkParams &P)
{
    const int n = t->view.n, s = P.s;
    const size_t per_member = P.chain ? (size_t)(4 * n + 2) : (size_t)(2 * n + 1) + (size_t)s * n;
    const size_t bytes = per_member * 8 * G2_WARPS;
    const unsigned grid = (unsigned)((P.ld + G2_WARPS - 1) / G2_WARPS);
    if (t->view.rank == 5) {
        set_smem(rk_g2_kernel<5>, bytes);
        rk_g2_kernel<5><<<grid, G2_WARPS * 32, bytes, ctx().stream>>>(t->view, P);
    } else {
        set_smem(rk_g2_kernel<3>, bytes);
        rk_g2_kernel<3><<<grid, G2_WARPS * 32, bytes, ctx().stream>>>(t->view, P);
    }
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

static bool g3_enabled()
{
    const char *e = getenv("QGSB_RK_LARGE");     // "g2": force the warp-per-member kernel (A/B measurements)
    return !(e && !strcmp(e, "g2"));
}

// the host side of the plan, cached in the tensor handle
struct G3Host {
    G3Plan plan;
    int mb = 0, n_chunks = 0;
    bool pairs = false;   // two members per thread (rk_g3p_kernel)
    DevBuf<G3Entry> ent;
    DevBuf<int> meta;      // chunk list (4 ints each), then the padded row lengths
    size_t smem = 0;
};

}  // namespace qgsb

struct qgsb_tensor::G3Cache : qgsb::G3Host {};

namespace qgsb {

void g3_release(qgsb_tensor::G3Cache *c) { delete c; }

// rows -> groups of about equal (padded) entry counts -> chunks of whole rows
static bool g3_prepare(const qgsb_tensor *t)
{
    const int n = t->view.n;
    std::vector<int> rowlen(n);
    long total = 0;
    for (int i = 1; i <= n; ++i) {
        rowlen[i - 1] = (t->h_row_ptr[i + 1] - t->h_row_ptr[i] + 3) & ~3;
        if (rowlen[i - 1] > G3_CHUNK) return false;
        total += rowlen[i - 1];
    }
    if (total == 0) return false;
    auto *c = new qgsb_tensor::G3Cache();
    G3Plan &plan = c->plan;
    const char *mode = getenv("QGSB_G3_MODE");          // "single": one member per thread, four row groups (A/B)
    const bool pairs = !(mode && !strcmp(mode, "single"));
    const int G = (int)std::min<long>(pairs ? G3_MAX_GROUPS : 4, std::max<long>(1, total / (pairs ? 1024 : 2048)));
    plan.groups = G;
    std::vector<G3Chunk> chunks;
    int row = 0;
    long done = 0;
    int off = 0;
    for (int h = 0; h < G; ++h) {
        plan.chunk0[h] = (int)chunks.size();
        plan.row0[h] = row;
        const long target = total * (h + 1) / G;
        G3Chunk cur{off, 0, row, 0};
        while (row < n && (h == G - 1 || done + rowlen[row] / 2 < target)) {
            if (cur.count + rowlen[row] > G3_CHUNK) {
                chunks.push_back(cur);
                cur = G3Chunk{off, 0, row, 0};
            }
            cur.count += rowlen[row];
            cur.nrows += 1;
            off += rowlen[row];
            done += rowlen[row];
            ++row;
        }
        if (cur.nrows > 0 || chunks.size() == (size_t)plan.chunk0[h]) chunks.push_back(cur);
    }
    plan.chunk0[G] = (int)chunks.size();
    plan.row0[G] = n;
    c->n_chunks = (int)chunks.size();
    const size_t fixed = sizeof(G3Entry) * G3_STRIDE * G3_RING * G + sizeof(G3Chunk) * chunks.size() +
                         sizeof(int) * (size_t)((n + 3) & ~3);
    const size_t limit = ctx().smem_optin;
    if (fixed + 32 * 8 * (size_t)(n + 1) > limit) {
        delete c;
        return false;
    }
    // a row group is whole warps: 32 members per warp, or 64 with two members per thread
    int mb = (int)std::min<size_t>(128, (limit - fixed) / (8 * (size_t)(n + 1))) & ~31;
    if (pairs) mb &= ~63;
    if (mb < 32 || (pairs && mb < 64)) {
        delete c;
        return false;
    }
    c->pairs = pairs;
    c->mb = mb;
    c->smem = fixed + (size_t)(n + 1) * 8 * mb;
    std::vector<G3Entry> h;
    h.reserve(total);
    for (int i = 1; i <= n; ++i) {
        for (int e = t->h_row_ptr[i]; e < t->h_row_ptr[i + 1]; ++e) {
            const uint32_t j = t->h_ent[e].jk & 0xffffu;
            // two members per thread: flag the entries whose x_j is the one of the entry before in the same row
            const bool reuse = pairs && e > t->h_row_ptr[i] && (t->h_ent[e - 1].jk & 0xffffu) == j;
            h.push_back(G3Entry{t->h_ent[e].v, j * (uint32_t)mb * 8u | (reuse ? 1u : 0u),
                                (uint32_t)((t->h_ent[e].jk >> 16) * (uint32_t)mb * 8u)});
        }
        while (h.size() & 3) h.push_back(G3Entry{0., 0u, 0u});       // + 0 * x_0 * x_0
    }
    std::vector<int> meta(chunks.size() * 4 + rowlen.size());
    memcpy(meta.data(), chunks.data(), chunks.size() * sizeof(G3Chunk));
    memcpy(meta.data() + chunks.size() * 4, rowlen.data(), rowlen.size() * sizeof(int));
    c->ent.alloc(h.size());
    QGSB_CUDA(cudaMemcpy(c->ent.p, h.data(), h.size() * sizeof(G3Entry), cudaMemcpyHostToDevice));
    c->meta.alloc(meta.size());
    QGSB_CUDA(cudaMemcpy(c->meta.p, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
    t->g3_cache = c;
    return true;
}

static bool launch_g3(const qgsb_tensor *t, RkParams &P)
{
    const int n = t->view.n;
    if (t->g3_state < 0) return false;
    if (t->g3_state == 0) {
        t->g3_state = -1;
        if (!g3_prepare(t)) return false;
        t->g3_state = 1;
    }
    const qgsb_tensor::G3Cache *c = t->g3_cache;
    PoolBuf<double> d_acc((size_t)n * P.ld), d_xn((size_t)n * P.ld);
    const unsigned grid = (unsigned)((P.ld + c->mb - 1) / c->mb);
    if (c->pairs) {
        set_smem(rk_g3p_kernel, c->smem);
        rk_g3p_kernel<<<grid, (c->mb / 2) * c->plan.groups, c->smem, ctx().stream>>>(
            n, c->mb, c->n_chunks, c->plan, c->ent.p, reinterpret_cast<const G3Chunk *>(c->meta.p),
            c->meta.p + (size_t)c->n_chunks * 4, P, d_acc.p, d_xn.p);
    } else {
        set_smem(rk_g3_kernel, c->smem);
        rk_g3_kernel<<<grid, c->mb * c->plan.groups, c->smem, ctx().stream>>>(
            n, c->mb, c->n_chunks, c->plan, c->ent.p, reinterpret_cast<const G3Chunk *>(c->meta.p),
            c->meta.p + (size_t)c->n_chunks * 4, P, d_acc.p, d_xn.p);
    }
    count_launch();
    QGSB_CUDA(cudaGetLastError());
    // the scratch arrays go back to the pool here; every user of the pool runs on the same stream
    return true;
}

// members below which a block per member (thread per row) beats a thread per member; measured on B200 for
// MAOOAM-36: 0.5 us per step against 3 us, and a full wave of row blocks costs about one thread-per-member step
static long rows_threshold()
{
    const char *e = getenv("QGSB_RK_ROWS_MAX");     // A/B measurements: 0 disables the rows kernel
    return e ? atol(e) : 2048;
}

template <int RANK, int EFU>
static void launch_rows_efu(int n, const PackTables &pt, RkParams &P)
{
    const size_t bytes = 2 * (size_t)((n + 2) & ~1) * sizeof(double);
    const int threads = (n + 31) / 32 * 32;
    rk_rows_kernel<RANK, EFU><<<(unsigned)P.n_members, threads, bytes, ctx().stream>>>(n, pt.f_ent, pt.EF, P);
}

static bool launch_rows(const qgsb_tensor *t, RkParams &P)
{
    const int n = t->view.n;
    if (n > 256) return false;                      // larger bases: one thread per row is too coarse
    const PackTables &pt = pack_tables(t, false);
    if (pt.f_ent == nullptr || pt.EF > 64) return false;
    const bool r5 = t->view.rank == 5;
    if (pt.EF <= 8) r5 ? launch_rows_efu<5, 8>(n, pt, P) : launch_rows_efu<3, 8>(n, pt, P);
    else if (pt.EF <= 16) r5 ? launch_rows_efu<5, 16>(n, pt, P) : launch_rows_efu<3, 16>(n, pt, P);
    else if (pt.EF <= 32) r5 ? launch_rows_efu<5, 32>(n, pt, P) : launch_rows_efu<3, 32>(n, pt, P);
    else r5 ? launch_rows_efu<5, 64>(n, pt, P) : launch_rows_efu<3, 64>(n, pt, P);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
    return true;
}

// rows of a tensor dealt to WIDE_WARPS warps, longest first onto the least loaded warp
static bool launch_wide(const qgsb_tensor *t, RkParams &P)
{
    const int n = t->view.n;
    const size_t bytes = (2 * (size_t)((n + 2) & ~1) + 2 * (size_t)((n + 1) & ~1)) * sizeof(double) +
                         3 * (size_t)n * sizeof(int);
    if (bytes > ctx().smem_optin) return false;
    if (t->d_wide.n == 0) {
        std::vector<int> order(n), load(WIDE_WARPS, 0);
        std::iota(order.begin(), order.end(), 1);
        auto len = [&](int i) { return t->h_row_ptr[i + 1] - t->h_row_ptr[i]; };
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len(x) > len(y); });
        std::vector<std::vector<int>> rows(WIDE_WARPS);
        for (int i : order) {
            const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            rows[w].push_back(i);
            load[w] += len(i) + 8;                     // + the fixed cost of finishing a row
        }
        std::vector<int> flat(WIDE_WARPS + 1, 0);
        for (int w = 0; w < WIDE_WARPS; ++w) flat[w + 1] = flat[w] + (int)rows[w].size();
        for (int w = 0; w < WIDE_WARPS; ++w) flat.insert(flat.end(), rows[w].begin(), rows[w].end());
        t->d_wide.alloc(flat.size());
        QGSB_CUDA(cudaMemcpy(t->d_wide.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    const int *warp_ptr = t->d_wide.p, *warp_rows = t->d_wide.p + WIDE_WARPS + 1;
    if (t->view.rank == 5) {
        set_smem(rk_wide_kernel<5>, bytes);
        rk_wide_kernel<5><<<(unsigned)P.n_members, WIDE_WARPS * 32, bytes, ctx().stream>>>(t->view, warp_ptr, warp_rows, P);
    } else {
        set_smem(rk_wide_kernel<3>, bytes);
        rk_wide_kernel<3><<<(unsigned)P.n_members, WIDE_WARPS * 32, bytes, ctx().stream>>>(t->view, warp_ptr, warp_rows, P);
    }
    count_launch();
    QGSB_CUDA(cudaGetLastError());
    return true;
}

void rk_advance(const qgsb_tensor *t, double *d_y, long ld, long N, long n_steps, const double *d_dt,
                const Tableau &tab, long write_steps, long R, double *d_rec)
{
    if (tab.chain && N <= rows_threshold() && n_steps > 0) {
        RkParams P;
        fill_params(P, tab, d_y, ld, N, n_steps, d_dt, write_steps, R, d_rec);
        if (launch_rows(t, P)) return;
        // many entries per row (large bases, T4): a block of 32 warps per member
        if (t->view.nnz >= 4096 && launch_wide(t, P)) return;
    }
    if (t->spec && t->use_spec && tab.chain && tab.s <= 8 && t->spec->rk_chain) {
        QGSB_CUDA(t->spec->rk_chain(d_y, ld, N, n_steps, d_dt, tab.s, tab.alpha.data(), tab.b.data(), write_steps, R,
                                    d_rec, ctx().sm_count, ctx().stream));
        count_launch();
        return;
    }
    if (t->spec && t->use_spec && !tab.chain && tab.s <= 8 && t->spec->rk_general) {
        // any other explicit tableau on the tensor-specialised code (stage derivatives in shared memory)
        const cudaError_t e = t->spec->rk_general(d_y, ld, N, n_steps, d_dt, tab.s, tab.a.data(), tab.b.data(),
                                                  write_steps, R, d_rec, ctx().smem_optin, ctx().stream);
        if (e == cudaSuccess) {
            count_launch();
            return;
        }
        QGSB_REQUIRE(e == cudaErrorInvalidValue, "specialised Runge-Kutta launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();                 // too many stages for shared memory: the generic kernel below serves
    }
    RkParams P;
    fill_params(P, tab, d_y, ld, N, n_steps, d_dt, write_steps, R, d_rec);
    const int n = t->view.n;
    if (n > QGSB_G1_MAX_NDIM && t->view.rank == 3 && tab.chain && g3_enabled() && launch_g3(t, P)) return;
    if (n <= QGSB_G1_MAX_NDIM) {
        const size_t per_thread = tab.chain ? (size_t)(4 * n + 2) : (size_t)(2 * n + 1) + (size_t)tab.s * n;
        if (per_thread * 8 * 64 + 4096 <= ctx().smem_optin) return launch_g1<64>(t, P);
        if (per_thread * 8 * 32 + 4096 <= ctx().smem_optin) return launch_g1<32>(t, P);
    }
    launch_g2(t, P);
}

// ------------------------------------------------------------------------------------------------
// DFMA peak micro-benchmark: 16 independent chains per thread, no memory traffic
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = seed + q + threadIdx.x * 1e-9;
    const double m = 0.999999, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = fma(a[q], m, c);
    }
    double ssum = 0.;
#pragma unroll
    for (int q = 0; q < 16; ++q) ssum += a[q];
    if (ssum == 12345.678) out[0] = ssum;  // never true: keeps the chains alive
}

// DMMA (mma.sync m8n8k4 f64) peak micro-benchmark: 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed)
{
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        c[q][0] = seed + q;
        c[q][1] = seed - q;
    }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 0.999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[q][0]), "+d"(c[q][1])
                         : "d"(a), "d"(b));
    }
    double ssum = 0.;
#pragma unroll
    for (int q = 0; q < 8; ++q) ssum += c[q][0] + c[q][1];
    if (ssum == 12345.678) out[0] = ssum;
}

}  // namespace qgsb

using namespace qgsb;

static long records_for(long n_steps, long write_steps)
{
    // integrate.py:190-196 with L = n_steps + 1 time points
    if (write_steps == 0) return 1;
    const long L = n_steps + 1;
    long r = (L + write_steps - 1) / write_steps;
    if ((r - 1) * write_steps != L - 1) r += 1;
    return r;
}

extern "C" {

// Trajectory streaming (the chunked-run idiom of qgs_maooam.py:115-136 with
// write_steps > 0): the records of a bounded number of write steps are kept in HBM, turned into the API layout and
// copied into the strided slice traj[:, :, r0:r1] of the caller's (N, n, R) array on the copy stream while the next
// chunk integrates (two buffers).  Device memory does not grow with R.  d_y (n, ld) tiled is advanced in place.
// ev0 / ev1 of the context bracket the integration.
static void stream_trajectories(const qgsb_tensor *t, double *d_y, long ld, long N, long n_steps, const double *d_dt,
                                const Tableau &tab, long write_steps, int time_direction, long R, double *traj)
{
    Context &cx = ctx();
    cudaStream_t st = cx.stream, so = cx.copy_out;
    const int n = t->view.n;
    const long rows = N * n;
    const bool flip = time_direction == -1;
    // regular records r = 0 .. R-2 (state before step r * write_steps) in chunks, then the final state as record R-1
    const size_t rec_bytes = (size_t)n * ld * sizeof(double);
    const long budget = std::max<long>(2, (long)(std::min<size_t>(cx.total_mem / 16, (size_t)4 << 30) / rec_bytes));
    long per_chunk = std::max<long>(1, std::min<long>(std::max<long>(R - 1, 1), budget - 1));
    if (const char *env = getenv("QGSB_STREAM_RECORDS")) per_chunk = std::max<long>(1, std::min<long>(per_chunk, atol(env)));
    PoolBuf<double> d_rec((size_t)(per_chunk + 1) * n * ld);
    PoolBuf<double> d_out0((size_t)rows * per_chunk), d_out1((size_t)rows * per_chunk);
    cudaEvent_t ready[2], freed[2];
    for (int q = 0; q < 2; ++q) {
        QGSB_CUDA(cudaEventCreateWithFlags(&ready[q], cudaEventDisableTiming));
        QGSB_CUDA(cudaEventCreateWithFlags(&freed[q], cudaEventDisableTiming));
    }
    QGSB_CUDA(cudaEventRecord(cx.ev0, st));
    // A chunk is shipped AFTER the next one has been launched: with a pageable destination the copy blocks the host
    // thread until the chunk has arrived, and in this order it does so while the GPU integrates the next chunk (with a
    // pinned destination everything is asynchronous and the order does not matter).
    struct Pending {
        double *d_out;
        long r0, count;
        int buf;
        bool valid;
    } pending = {nullptr, 0, 0, 0, false};
    bool used[2] = {false, false};
    auto ship = [&](const Pending &p) {
        // host columns [h0, h0 + count) of every (member, variable) row; the record axis is reversed for backward runs
        const long h0 = flip ? R - (p.r0 + p.count) : p.r0;
        QGSB_CUDA(cudaStreamWaitEvent(so, ready[p.buf], 0));
        QGSB_CUDA(cudaMemcpy2DAsync(traj + h0, (size_t)R * sizeof(double), p.d_out, (size_t)p.count * sizeof(double),
                                    (size_t)p.count * sizeof(double), (size_t)rows, cudaMemcpyDeviceToHost, so));
        QGSB_CUDA(cudaEventRecord(freed[p.buf], so));
    };
    int buf = 0;
    auto produce = [&](long r0, long count, const double *d_src, long src_records, int src_flip) {
        double *d_out = buf ? d_out1.p : d_out0.p;
        if (used[buf]) QGSB_CUDA(cudaStreamWaitEvent(st, freed[buf], 0));      // the buffer's last shipment has left
        launch_rec_to_api(d_src, d_out, N, n, src_records, ld, src_flip);
        QGSB_CUDA(cudaEventRecord(ready[buf], st));
        used[buf] = true;
        const Pending mine = {d_out, r0, count, buf, true};
        buf ^= 1;
        return mine;
    };
    if (write_steps > 0) {
        for (long r0 = 0; r0 < R - 1; r0 += per_chunk) {
            const long r1 = std::min(R - 1, r0 + per_chunk);
            const long step0 = r0 * write_steps, step1 = std::min(n_steps, r1 * write_steps);
            const long steps = step1 - step0, rc = records_for(steps, write_steps);
            rk_advance(t, d_y, ld, N, steps, d_dt + step0, tab, write_steps, rc, d_rec.p);
            const Pending mine = produce(r0, r1 - r0, d_rec.p, r1 - r0, flip ? 1 : 0);
            if (pending.valid) ship(pending);
            pending = mine;
        }
    } else {
        rk_advance(t, d_y, ld, N, n_steps, d_dt, tab, 0, 1, nullptr);
    }
    {
        const Pending last = produce(R - 1, 1, d_y, 1, 0);
        QGSB_CUDA(cudaEventRecord(cx.ev1, st));
        if (pending.valid) ship(pending);
        ship(last);
    }
    QGSB_CUDA(cudaStreamSynchronize(so));
    QGSB_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < 2; ++q) {
        cudaEventDestroy(ready[q]);
        cudaEventDestroy(freed[q]);
    }
}

// one device's share of qgsb_rk_integrate: members [0, N) of ic / traj on the calling thread's device
static void rk_integrate_device(const qgsb_tensor *t, long N, const double *ic, long n_steps, const double *dt, int s,
                                const double *a, const double *b, long write_steps, int time_direction, long R,
                                double *traj, double *device_ms)
{
    Context &cx = ctx();
    cudaStream_t st = cx.stream;
    const Tableau tab = make_tableau(s, a, b);
    const int n = t->view.n;
    const long ld = round_up(N, TILE);
    PoolBuf<double> d_ic((size_t)N * n), d_y((size_t)n * ld), d_dt(std::max<long>(n_steps, 1));
    PoolBuf<double> d_out(R == 1 ? (size_t)N * n : 1);
    if (n_steps) d_dt.upload(dt, n_steps, st);
    if (R == 1) {
        // write_steps == 0 (or a single time point): only the end state is returned (integrate.py:221).
        // Large ensembles go through in chunks of whole waves of thread blocks so that the upload of chunk k+1
        // and the download of chunk k-1 overlap the integration of chunk k (two copy streams, pinned host
        // buffers make the copies truly asynchronous; members are independent, so the results are unchanged).
        const long chunk = (long)cx.sm_count * 2 * TILE * 4;
        if (N >= 2 * chunk) {
            cudaStream_t s_in = cx.copy_in, s_out = cx.copy_out;
            const long n_chunks = (N + chunk - 1) / chunk;
            std::vector<cudaEvent_t> up(n_chunks), done(n_chunks);
            for (long k = 0; k < n_chunks; ++k) {
                QGSB_CUDA(cudaEventCreateWithFlags(&up[k], cudaEventDisableTiming));
                QGSB_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
            }
            QGSB_CUDA(cudaStreamSynchronize(st));      // d_dt upload and earlier work on the pool buffers
            // Issue order: compute(k) is launched BEFORE upload(k + 1) and download(k - 1) are issued.  With pinned
            // host buffers every call below is asynchronous and the three streams overlap by themselves; with ordinary
            // (pageable) numpy arrays a copy blocks the host thread -- an upload until the data is staged, a download
            // until it has arrived -- and in this order it blocks while the GPU is busy with chunk k, so the copies of
            // the neighbouring chunks still hide behind the integration.
            auto upload = [&](long k) {
                const long m0 = k * chunk, nk = std::min(chunk, N - m0);
                QGSB_CUDA(cudaMemcpyAsync(d_ic.p + (size_t)m0 * n, ic + (size_t)m0 * n, sizeof(double) * nk * n,
                                          cudaMemcpyHostToDevice, s_in));
                QGSB_CUDA(cudaEventRecord(up[k], s_in));
            };
            auto download = [&](long k) {
                const long m0 = k * chunk, nk = std::min(chunk, N - m0);
                QGSB_CUDA(cudaStreamWaitEvent(s_out, done[k], 0));
                QGSB_CUDA(cudaMemcpyAsync(traj + (size_t)m0 * n, d_out.p + (size_t)m0 * n, sizeof(double) * nk * n,
                                          cudaMemcpyDeviceToHost, s_out));
            };
            upload(0);
            QGSB_CUDA(cudaEventRecord(cx.ev0, st));
            for (long k = 0; k < n_chunks; ++k) {
                const long m0 = k * chunk, nk = std::min(chunk, N - m0), ldk = round_up(nk, TILE);
                double *yk = d_y.p + (size_t)m0 * n;   // chunk starts on a tile boundary: its tiles are contiguous
                QGSB_CUDA(cudaStreamWaitEvent(st, up[k], 0));
                launch_aos_to_soa(d_ic.p + (size_t)m0 * n, yk, nk, n, ldk);
                rk_advance(t, yk, ldk, nk, n_steps, d_dt.p, tab, 0, 1, nullptr);
                launch_soa_to_aos(yk, d_out.p + (size_t)m0 * n, nk, n, ldk);
                QGSB_CUDA(cudaEventRecord(done[k], st));
                if (k + 1 < n_chunks) upload(k + 1);
                if (k >= 1) download(k - 1);
            }
            download(n_chunks - 1);
            QGSB_CUDA(cudaEventRecord(cx.ev1, st));
            QGSB_CUDA(cudaStreamSynchronize(s_out));
            QGSB_CUDA(cudaStreamSynchronize(st));
            for (long k = 0; k < n_chunks; ++k) {
                cudaEventDestroy(up[k]);
                cudaEventDestroy(done[k]);
            }
            if (device_ms) {
                float ms = 0.f;
                QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
                *device_ms = ms;
            }
            return;
        }
        d_ic.upload(ic, (size_t)N * n, st);
        launch_aos_to_soa(d_ic.p, d_y.p, N, n, ld);
        QGSB_CUDA(cudaEventRecord(cx.ev0, st));
        rk_advance(t, d_y.p, ld, N, n_steps, d_dt.p, tab, 0, 1, nullptr);
        QGSB_CUDA(cudaEventRecord(cx.ev1, st));
        launch_soa_to_aos(d_y.p, d_out.p, N, n, ld);
    } else {
        // trajectories: records stream to the caller's array chunk by chunk (bounded device memory)
        d_ic.upload(ic, (size_t)N * n, st);
        launch_aos_to_soa(d_ic.p, d_y.p, N, n, ld);
        stream_trajectories(t, d_y.p, ld, N, n_steps, d_dt.p, tab, write_steps, time_direction, R, traj);
        if (device_ms) {
            float ms = 0.f;
            QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
            *device_ms = ms;
        }
        return;
    }
    d_out.download(traj, (size_t)N * n * R, st);
    QGSB_CUDA(cudaStreamSynchronize(st));
    if (device_ms) {
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        *device_ms = ms;
    }
}

// Members are independent ODE solves: with several devices the ensemble is split into contiguous blocks, one per
// device, each integrated by the single-device path above on its own streams (the reference deals trajectories to its
// worker processes, integrator.py:386-395).  A shard never falls under the size at which the whole ensemble would
// have chosen another kernel family, so the results are bitwise those of one device.
int qgsb_rk_integrate(const qgsb_tensor *t, long N, const double *ic, long n_steps, const double *dt, int s,
                      const double *a, const double *b, const double *c, long write_steps, int time_direction,
                      long R, double *traj, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && ic && traj, "null argument");
    QGSB_REQUIRE(N >= 1, "need at least one trajectory, got %ld", N);
    QGSB_REQUIRE(n_steps >= 0 && write_steps >= 0, "negative step count");
    QGSB_REQUIRE(n_steps == 0 || dt != nullptr, "null dt array");
    QGSB_REQUIRE(time_direction == 1 || time_direction == -1, "time_direction must be +1 or -1");
    QGSB_REQUIRE(R == records_for(n_steps, write_steps), "n_records %ld inconsistent with %ld steps / write_steps %ld",
                 R, n_steps, write_steps);
    ensure_init();
    const int n = t->view.n;
    const int parts = shard_count(N, std::max<long>(4096, 2 * rows_threshold()));
    std::vector<double> ms(parts, 0.);
    run_sharded(N, parts, [&](int g, long lo, long hi) {
        rk_integrate_device(tensor_here(t), hi - lo, ic + (size_t)lo * n, n_steps, dt, s, a, b, write_steps,
                            time_direction, R, traj + (size_t)lo * n * R, device_ms ? &ms[g] : nullptr);
    });
    if (device_ms) *device_ms = *std::max_element(ms.begin(), ms.end());
    QGSB_API_END
}

// ---- resident ensemble ---------------------------------------------------------------------------
// Every entry point has a single-device implementation (the code of round 1, on the calling thread's device) and a
// wrapper that runs it directly or, for a composite ensemble, once per part on the part's device.
}  // extern "C"

qgsb_ensemble::~qgsb_ensemble()
{
    for (qgsb_ensemble *p : parts) {
        cudaSetDevice(p->device);
        delete p;
    }
    if (!parts.empty() && qgsb::device_slots() > 0) cudaSetDevice(qgsb::ctx().device);
}

static qgsb_ensemble *ensemble_create_device(const qgsb_tensor *t, long N)
{
    qgsb_ensemble *e = new qgsb_ensemble();
    try {
        e->tensor = t;
        e->N = N;
        e->ld = round_up(N, TILE);
        e->device = ctx().device;
        e->d_y.alloc((size_t)t->view.n * e->ld);
        e->d_stage.alloc((size_t)N * t->view.n);
        QGSB_CUDA(cudaMemsetAsync(e->d_y.p, 0, sizeof(double) * t->view.n * e->ld, ctx().stream));
    } catch (...) {
        delete e;
        throw;
    }
    return e;
}

static void ensemble_upload_device(qgsb_ensemble *e, const double *ic)
{
    const int n = e->tensor->view.n;
    e->d_stage.upload(ic, (size_t)e->N * n, ctx().stream);
    launch_aos_to_soa(e->d_stage.p, e->d_y.p, e->N, n, e->ld);
}

static void ensemble_download_device(qgsb_ensemble *e, double *out)
{
    const int n = e->tensor->view.n;
    launch_soa_to_aos(e->d_y.p, e->d_stage.p, e->N, n, e->ld);
    e->d_stage.download(out, (size_t)e->N * n, ctx().stream);
    QGSB_CUDA(cudaStreamSynchronize(ctx().stream));
}

static void ensemble_run(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a, const double *b,
                         long write_steps, long R, double *d_rec, double *device_ms)
{
    QGSB_REQUIRE(e, "null ensemble");
    QGSB_REQUIRE(n_steps >= 0, "negative step count");
    Context &cx = ctx();
    const Tableau tab = make_tableau(s, a, b);
    // dt staging buffer lives with the ensemble so the launch can stay asynchronous
    DevBuf<double> &d_dt = e->d_dt;
    if (d_dt.n < (size_t)std::max<long>(n_steps, 1)) {
        QGSB_CUDA(cudaStreamSynchronize(cx.stream));
        d_dt.alloc(std::max<long>(n_steps, 1));
    }
    if (n_steps) d_dt.upload(dt, n_steps, cx.stream);
    if (device_ms) QGSB_CUDA(cudaEventRecord(cx.ev0, cx.stream));
    rk_advance(e->tensor, e->d_y.p, e->ld, e->N, n_steps, d_dt.p, tab, write_steps, R, d_rec);
    if (device_ms) {
        QGSB_CUDA(cudaEventRecord(cx.ev1, cx.stream));
        QGSB_CUDA(cudaEventSynchronize(cx.ev1));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        *device_ms = ms;
    }
}

// Ensemble statistics without the (N, n, R) dump of TrajectoriesStatistics.compute_stats
// (qgs/integrators/statistics.py:33-66): integrate, keep the records of a bounded number of write steps in HBM,
// reduce them to per-record sums on the device, continue.  Record semantics are those of integrate.py:190-221.
static void ensemble_moments_run_device(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                        const double *b, long write_steps, long R, double *sum, double *sumsq,
                                        double *device_ms)
{
    Context &cx = ctx();
    cudaStream_t st = cx.stream;
    const Tableau tab = make_tableau(s, a, b);
    const int n = e->tensor->view.n;
    const long ld = e->ld;
    PoolBuf<double> d_dt(std::max<long>(n_steps, 1)), d_sum((size_t)R * n), d_sq((size_t)R * n);
    if (n_steps) d_dt.upload(dt, n_steps, st);
    QGSB_CUDA(cudaEventRecord(cx.ev0, st));
    if (write_steps == 0 || R == 1) {
        rk_advance(e->tensor, e->d_y.p, ld, e->N, n_steps, d_dt.p, tab, 0, 1, nullptr);
        launch_record_moments(e->d_y.p, 1, e->N, n, ld, d_sum.p, d_sq.p);
    } else {
        // regular records r = 0 .. R-2 are the states before step r * write_steps; record R-1 is the final state
        const size_t rec_bytes = (size_t)n * ld * sizeof(double);
        const long budget = std::max<long>(2, (long)(std::min<size_t>(cx.total_mem / 8, (size_t)8 << 30) / rec_bytes));
        const long per_chunk = std::min<long>(R - 1, budget - 1);
        PoolBuf<double> d_rec((size_t)(per_chunk + 1) * n * ld);
        for (long r0 = 0; r0 < R - 1; r0 += per_chunk) {
            const long r1 = std::min(R - 1, r0 + per_chunk);
            const long step0 = r0 * write_steps, step1 = std::min(n_steps, r1 * write_steps);
            const long steps = step1 - step0, rc = records_for(steps, write_steps);
            rk_advance(e->tensor, e->d_y.p, ld, e->N, steps, d_dt.p + step0, tab, write_steps, rc, d_rec.p);
            launch_record_moments(d_rec.p, r1 - r0, e->N, n, ld, d_sum.p + (size_t)r0 * n, d_sq.p + (size_t)r0 * n);
        }
        launch_record_moments(e->d_y.p, 1, e->N, n, ld, d_sum.p + (size_t)(R - 1) * n, d_sq.p + (size_t)(R - 1) * n);
    }
    QGSB_CUDA(cudaEventRecord(cx.ev1, st));
    d_sum.download(sum, (size_t)R * n, st);
    d_sq.download(sumsq, (size_t)R * n, st);
    QGSB_CUDA(cudaStreamSynchronize(st));
    if (device_ms) {
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        *device_ms = ms;
    }
}

static void ensemble_trajectories_device(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                         const double *b, long write_steps, int time_direction, long R, double *traj,
                                         double *device_ms)
{
    Context &cx = ctx();
    const Tableau tab = make_tableau(s, a, b);
    PoolBuf<double> d_dt(std::max<long>(n_steps, 1));
    if (n_steps) d_dt.upload(dt, n_steps, cx.stream);
    stream_trajectories(e->tensor, e->d_y.p, e->ld, e->N, n_steps, d_dt.p, tab, write_steps, time_direction, R, traj);
    if (device_ms) {
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        *device_ms = ms;
    }
}

// runs body(g, part, first member of the part) for every part of a composite on the part's device; returns the largest
// of the device times the bodies report
template <class F>
static double over_parts(qgsb_ensemble *e, F &&body)
{
    const int parts = (int)e->parts.size();
    std::vector<double> ms(parts, 0.);
    run_sharded(e->N, parts, [&](int g, long, long) { ms[g] = body(g, e->parts[g], e->lo[g]); });
    return *std::max_element(ms.begin(), ms.end());
}

extern "C" {

int qgsb_ensemble_create(const qgsb_tensor *t, long N, qgsb_ensemble **out)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && out, "null argument");
    QGSB_REQUIRE(N >= 1, "need at least one trajectory");
    ensure_init();
    const int parts = shard_count(N, std::max<long>(4096, 2 * rows_threshold()));
    if (parts <= 1) {
        *out = ensemble_create_device(t, N);
        return 0;
    }
    qgsb_ensemble *e = new qgsb_ensemble();
    e->tensor = t;
    e->N = N;
    e->device = ctx().device;
    e->parts.assign(parts, nullptr);
    e->lo.assign(parts + 1, 0);
    for (int g = 0; g <= parts; ++g) e->lo[g] = (long)((__int128)g * N / parts);      // shard_range of run_sharded
    try {
        run_sharded(N, parts, [&](int g, long lo, long hi) { e->parts[g] = ensemble_create_device(tensor_here(t), hi - lo); });
    } catch (...) {
        delete e;
        throw;
    }
    *out = e;
    QGSB_API_END
}

void qgsb_ensemble_destroy(qgsb_ensemble *e)
{
    if (!e) return;
    QGSB_API_LOCK
    if (device_slots() > 0 && ctx().ready) {
        cudaSetDevice(ctx().device);
        cudaStreamSynchronize(ctx().stream);
    }
    delete e;
}

int qgsb_ensemble_upload(qgsb_ensemble *e, const double *ic)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(e && ic, "null argument");
    ensure_init();
    const int n = e->tensor->view.n;
    if (e->parts.empty()) {
        ensemble_upload_device(e, ic);
    } else {
        over_parts(e, [&](int, qgsb_ensemble *p, long lo) {
            ensemble_upload_device(p, ic + (size_t)lo * n);
            QGSB_CUDA(cudaStreamSynchronize(ctx().stream));     // the worker thread's stream must be done before it leaves
            return 0.;
        });
    }
    QGSB_API_END
}

int qgsb_ensemble_download(qgsb_ensemble *e, double *out)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(e && out, "null argument");
    ensure_init();
    const int n = e->tensor->view.n;
    if (e->parts.empty()) {
        ensemble_download_device(e, out);
    } else {
        over_parts(e, [&](int, qgsb_ensemble *p, long lo) {
            ensemble_download_device(p, out + (size_t)lo * n);
            return 0.;
        });
    }
    QGSB_API_END
}

int qgsb_ensemble_integrate(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                            const double *b, const double *c, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(e, "null ensemble");
    ensure_init();
    if (e->parts.empty()) {
        ensemble_run(e, n_steps, dt, s, a, b, 0, 1, nullptr, device_ms);
    } else {
        // a composite is always integrated to completion: the worker threads do not outlive the call
        const double ms = over_parts(e, [&](int, qgsb_ensemble *p, long) {
            double part = 0.;
            ensemble_run(p, n_steps, dt, s, a, b, 0, 1, nullptr, &part);
            return part;
        });
        if (device_ms) *device_ms = ms;
    }
    QGSB_API_END
}

int qgsb_ensemble_integrate_record(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                   const double *b, const double *c, long write_steps, long R, double *d_rec,
                                   double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(e, "null ensemble");
    QGSB_REQUIRE(d_rec != nullptr, "null record buffer");
    QGSB_REQUIRE(e->parts.empty(), "a record buffer on one device cannot take an ensemble spread over %zu devices; use "
                 "qgsb_ensemble_integrate_trajectories or create the ensemble under qgsb_set_devices(1, ...)",
                 e->parts.size());
    QGSB_REQUIRE(R == records_for(n_steps, write_steps), "n_records %ld inconsistent with %ld steps / write_steps %ld",
                 R, n_steps, write_steps);
    ensure_init();
    ensemble_run(e, n_steps, dt, s, a, b, write_steps, R, d_rec, device_ms);
    QGSB_API_END
}

int qgsb_ensemble_integrate_moments(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                    const double *b, const double *c, long write_steps, long R, double *sum,
                                    double *sumsq, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(e && sum && sumsq, "null argument");
    QGSB_REQUIRE(n_steps >= 0 && write_steps >= 0, "negative step count");
    QGSB_REQUIRE(n_steps == 0 || dt != nullptr, "null dt array");
    QGSB_REQUIRE(R == records_for(n_steps, write_steps), "n_records %ld inconsistent with %ld steps / write_steps %ld",
                 R, n_steps, write_steps);
    ensure_init();
    if (e->parts.empty()) {
        ensemble_moments_run_device(e, n_steps, dt, s, a, b, write_steps, R, sum, sumsq, device_ms);
    } else {
        // per-part sums, added in part order (deterministic)
        const int n = e->tensor->view.n;
        const size_t len = (size_t)R * n;
        const int parts = (int)e->parts.size();
        std::vector<double> s1(len * parts), s2(len * parts);
        const double ms = over_parts(e, [&](int g, qgsb_ensemble *p, long) {
            double part = 0.;
            ensemble_moments_run_device(p, n_steps, dt, s, a, b, write_steps, R, s1.data() + len * g,
                                        s2.data() + len * g, &part);
            return part;
        });
        for (size_t q = 0; q < len; ++q) {
            double t1 = 0., t2 = 0.;
            for (int g = 0; g < parts; ++g) {
                t1 += s1[len * g + q];
                t2 += s2[len * g + q];
            }
            sum[q] = t1;
            sumsq[q] = t2;
        }
        if (device_ms) *device_ms = ms;
    }
    QGSB_API_END
}

// The resident ensemble's records streamed to the host (see stream_trajectories).
int qgsb_ensemble_integrate_trajectories(qgsb_ensemble *e, long n_steps, const double *dt, int s, const double *a,
                                         const double *b, const double *c, long write_steps, int time_direction,
                                         long R, double *traj, double *device_ms)
{
    (void)c;
    QGSB_API_BEGIN
    QGSB_REQUIRE(e && traj, "null argument");
    QGSB_REQUIRE(n_steps >= 0 && write_steps >= 0, "negative step count");
    QGSB_REQUIRE(n_steps == 0 || dt != nullptr, "null dt array");
    QGSB_REQUIRE(time_direction == 1 || time_direction == -1, "time_direction must be +1 or -1");
    QGSB_REQUIRE(R == records_for(n_steps, write_steps), "n_records %ld inconsistent with %ld steps / write_steps %ld",
                 R, n_steps, write_steps);
    ensure_init();
    if (e->parts.empty()) {
        ensemble_trajectories_device(e, n_steps, dt, s, a, b, write_steps, time_direction, R, traj, device_ms);
    } else {
        const int n = e->tensor->view.n;
        const double ms = over_parts(e, [&](int, qgsb_ensemble *p, long lo) {
            double part = 0.;
            ensemble_trajectories_device(p, n_steps, dt, s, a, b, write_steps, time_direction, R,
                                         traj + (size_t)lo * n * R, &part);
            return part;
        });
        if (device_ms) *device_ms = ms;
    }
    QGSB_API_END
}

// Device pointer / leading dimension of the state: defined for an ensemble on ONE device only (a composite returns
// NULL / 0).
void *qgsb_ensemble_device_ptr(qgsb_ensemble *e) { return (e && e->parts.empty()) ? (void *)e->d_y.p : nullptr; }
long qgsb_ensemble_ld(const qgsb_ensemble *e) { return (e && e->parts.empty()) ? e->ld : 0; }

int qgsb_dmma_peak(double *tflops)
{
    QGSB_API_BEGIN
    ensure_init();
    Context &cx = ctx();
    DevBuf<double> d_out(1);
    const int iters = 2048, blocks = cx.sm_count * 8, threads = 256;
    dmma_peak_kernel<<<blocks, threads, 0, cx.stream>>>(d_out.p, 16, 1.0);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        QGSB_CUDA(cudaEventRecord(cx.ev0, cx.stream));
        dmma_peak_kernel<<<blocks, threads, 0, cx.stream>>>(d_out.p, iters, 1.0 + rep);
        QGSB_CUDA(cudaEventRecord(cx.ev1, cx.stream));
        QGSB_CUDA(cudaEventSynchronize(cx.ev1));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        best = std::min(best, ms);
    }
    count_launch(6);
    // one m8n8k4 mma = 8*8*4 FMA per warp
    const double flops = 2.0 * 256.0 * 8.0 * (double)iters * (double)blocks * (threads / 32);
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    QGSB_API_END
}

int qgsb_fp64_peak(double *tflops, double *ms_out)
{
    QGSB_API_BEGIN
    ensure_init();
    Context &cx = ctx();
    DevBuf<double> d_out(1);
    const int iters = 4096, blocks = cx.sm_count * 8, threads = 256;
    dfma_peak_kernel<<<blocks, threads, 0, cx.stream>>>(d_out.p, 64, 1.0);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        QGSB_CUDA(cudaEventRecord(cx.ev0, cx.stream));
        dfma_peak_kernel<<<blocks, threads, 0, cx.stream>>>(d_out.p, iters, 1.0 + rep);
        QGSB_CUDA(cudaEventRecord(cx.ev1, cx.stream));
        QGSB_CUDA(cudaEventSynchronize(cx.ev1));
        float ms = 0.f;
        QGSB_CUDA(cudaEventElapsedTime(&ms, cx.ev0, cx.ev1));
        best = std::min(best, ms);
    }
    count_launch(6);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    QGSB_API_END
}

}  // extern "C"
