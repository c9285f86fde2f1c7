// tgls_shared.cuh -- declarations shared by the tangent-linear kernels (tgls.cu, tgls_pack.cuh, clv.cu).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace qgsb {

constexpr int TG_THREADS = 128;

struct TgParams {
    long n_members;
    int m;                 // tangent columns
    int s;
    int adjoint;
    double inverse;
    double a[16 * 16];
    double b[16];
    // --- plain TGLS integration (integrate.py:555-614) ---
    long n_steps;
    const double *dt;      // (n_steps)
    long write_steps;
    long n_records;
    double *y;             // (N, n)     in: ic, out: end state
    double *fm;            // (N, n, m)  in: tg_ic, out: end state
    double *rec_y;         // (R, N, n) or null
    double *rec_fm;        // (R, N, n, m) or null
    // --- Benettin (lyapunov.py:471-632) ---
    int forward;
    long n_pre, n_rec;
    const double *dt_macro;   // (n_pre + n_rec)
    const long *sub_ptr;      // (n_pre + n_rec + 1)
    const double *sub_dt;
    const double *stored;     // forward mode: write_steps=1 trajectory, tiled SoA records; else null
    long stored_ld;
    const long *start_idx;    // forward mode: stored-trajectory index used by every step
    long final_idx;
    const double *r0;         // (N, m, m) or null
    double *rec_exp;          // (R, N, m)
    double *r_all;            // (N, steps - r_first, m, m) or null: R of every step >= r_first
    long r_first;
    double *q_all;            // (N, n_rec + 1, n, m) or null
    int qr_at_start;          // != 0: fm holds a raw start matrix; factorise it before the first step (lyapunov.py:592-593)
    // --- placement of the big matrices ---
    double *scratch;          // global scratch when shared memory is too small, else null
    size_t scratch_per_member;
    // --- large bases: J as a dense (jd_ld x jd_ld) matrix per member in the scratch, product on the FP64 tensor cores ---
    int jd_ld;                // 0: sparse product over the position list; else leading dimension (n rounded up to 8)
};

template <int RANK>
__device__ __forceinline__ double f_row(const TensorView &T, int i, const double *xs)
{
    double acc = 0.;
    for (int e = T.row_ptr[i]; e < T.row_ptr[i + 1]; ++e) {
        const Entry en = T.ent[e];
        double p = xs[en.jk & 0xffffu] * xs[en.jk >> 16];
        if (RANK == 5) p = p * xs[en.lm & 0xffffu] * xs[en.lm >> 16];
        acc += p * en.v;
    }
    return acc;
}

template <int RANK>
__device__ __forceinline__ double jac_pos(const JacView &J, int p, const double *xs)
{
    double acc = 0.;
    for (int e = J.pos_ptr[p]; e < J.pos_ptr[p + 1]; ++e) {
        const Entry en = J.ent[e];
        double q = xs[en.jk & 0xffffu];
        if (RANK == 5) q = q * xs[en.jk >> 16] * xs[en.lm];
        acc += q * en.v;
    }
    return acc;
}

// run-time rank dispatch (the register kernels are templated on ndim only)
__device__ __forceinline__ double f_row_rt(const TensorView &T, int i, const double *xs)
{
    return T.rank == 5 ? f_row<5>(T, i, xs) : f_row<3>(T, i, xs);
}

__device__ __forceinline__ double jac_pos_rt(const JacView &J, int rank, int p, const double *xs)
{
    return rank == 5 ? jac_pos<5>(J, p, xs) : jac_pos<3>(J, p, xs);
}

// ---- tables of the packed kernels (tgls_pack.cuh) -----------------------------------------------------------------
// ELL ("padded column") forms of the tendency rows and of the Jacobian positions: entry e of row r sits at
// f_ent[e * n + r], entry e of list position q at j_ent[e * npos + q], so the m threads of a member read
// consecutive 16-byte records and every thread runs the same trip count.  Padding records have v = 0 and
// multiply x_0 = 1.  A table is absent (null) when padding would cost more than 1.5x the real entries; the
// kernels then walk the CSR lists of TensorView.
struct __align__(16) PEnt {
    double v;
    unsigned short a, b, c, d;   // BYTE offsets (8 * index) of the factors in the augmented state (0 = the constant 1);
                                 // Jacobian tables: d of entry 0 = byte offset of the position's slot;
                                 // tendency tables of rank 3: d of entry 0 = number of entries of the row
};

struct PackTables {
    const PEnt *f_ent = nullptr;   // device, (EF, n)
    int EF = 0;
    const PEnt *j_ent = nullptr;   // device, (EJ, npos)
    int EJ = 0;
    int npos = 0;
};

// ELL tables of a tensor handle, built on first use and cached in it (tgls_pack.cu); spec: Jacobian positions in the
// slot order of the handle's generated module instead of the dense n x n layout
const PackTables &pack_tables(const qgsb_tensor *t, bool spec);

// Benettin plumbing shared with clv.cu (tgls.cu)
void benettin_dispatch(const qgsb_tensor *t, const Tableau &tab, TgParams &P, DevBuf<double> &scratch);
void benettin_fill_common(TgParams &P, const Tableau &tab, long N, int m, int adjoint, double inverse);
void launch_transpose_records(const double *d_in, double *d_out, long R, long inner, int flip);

// packed kernels (tgls_pack.cu / tgls_pack.cuh)
bool pack_tangent_supported(const qgsb_tensor *t, const Tableau &tab, int m);
void launch_pack_tangent(const qgsb_tensor *t, const TgParams &P, bool lyap);

}  // namespace qgsb
