// tgls_pack.cu -- dispatch of the packed tangent-linear / Benettin kernels (tgls_pack.cuh).
//
// A tensor with a generated module (spec_registry.h: SpecKernels::tangent) runs the product over its
// literal Jacobian position list; every other tensor of a supported ndim runs the dense n x n product.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "spec_registry.h"
#include "tgls_pack.cuh"

namespace qgsb {

static int forced_kernel()
{
    // QGSB_TGLS_KERNEL = pack | pack_dense | generic  (testing / A-B measurements)
    const char *e = getenv("QGSB_TGLS_KERNEL");
    if (!e || !e[0]) return 0;
    if (!strcmp(e, "pack")) return 1;
    if (!strcmp(e, "pack_dense")) return 2;
    if (!strcmp(e, "generic")) return 4;
    return 0;
}

// dense-product kernels of one ndim: each instantiation lives in its own translation unit (tgls_pack_dense_<n>.cu),
// so that the three of them compile side by side
cudaError_t launch_pack_dense_20(const TensorView &T, const TgParams &P, const PackTables &tab, bool lyap, size_t smem,
                                 cudaStream_t stream);
cudaError_t launch_pack_dense_36(const TensorView &T, const TgParams &P, const PackTables &tab, bool lyap, size_t smem,
                                 cudaStream_t stream);
cudaError_t launch_pack_dense_38(const TensorView &T, const TgParams &P, const PackTables &tab, bool lyap, size_t smem,
                                 cudaStream_t stream);

static bool dense_ndim(int n) { return n == 20 || n == 36 || n == 38; }

static bool spec_tangent_usable(const qgsb_tensor *t)
{
    return t->spec && t->use_spec && t->spec->tangent && t->jac_matches_spec;
}

bool pack_tangent_supported(const qgsb_tensor *t, const Tableau &tab, int m)
{
    const int force = forced_kernel();
    if (force == 4) return false;
    if (!tab.chain || m < 1 || m > pack::MAX_THREADS) return false;
    const int n = t->view.n;
    if (!(spec_tangent_usable(t) && force != 2) && !dense_ndim(n)) return false;
    // a block must hold enough columns to fill its warps: with very few columns per member the per-member shared
    // state limits the packing and the generic kernel (one block per member) is the better fit
    const int jv = (spec_tangent_usable(t) && force != 2) ? t->spec->jac_slots : n * n;
    const size_t per_member = (size_t)(jv + 10 * n + 2 * (size_t)n * m + 3 * m + 40) * sizeof(double);
    const int G = std::min<int>(pack::MAX_THREADS / m, (int)(ctx().smem_optin / per_member));
    return G >= 1 && G * m >= 96;
}

}  // namespace qgsb

// device copies of the ELL tables of one tensor handle
struct qgsb_tensor::PackCache {
    qgsb::DevBuf<qgsb::PEnt> f, j;
    qgsb::PackTables tab;
};

qgsb_tensor::~qgsb_tensor()
{
    // replicas first: each frees its arrays on its own device
    for (auto &kv : replicas) {
        cudaSetDevice(kv.first);
        delete kv.second;
    }
    replicas.clear();
    cudaSetDevice(device);
    delete pack_cache[0];
    delete pack_cache[1];
    qgsb::g3_release(g3_cache);
}

namespace qgsb {

// a list is kept in CSR form when its ELL form would be long AND mostly padding
static bool ell_ok(int width, long count, long real)
{
    return width > 0 && (width <= 4 || (double)width * (double)count <= 1.5 * (double)real + 64.);
}

const PackTables &pack_tables(const qgsb_tensor *t, bool spec)
{
    qgsb_tensor::PackCache *&slot = t->pack_cache[spec ? 1 : 0];
    if (slot) return slot->tab;
    auto *pc = new qgsb_tensor::PackCache();
    const int n = t->view.n, rank = t->view.rank;
    // tendency rows 1..n
    {
        int EF = 0;
        long real = 0;
        for (int r = 1; r <= n; ++r) {
            const int len = t->h_row_ptr[r + 1] - t->h_row_ptr[r];
            EF = std::max(EF, len);
            real += len;
        }
        if (ell_ok(EF, n, real)) {
            std::vector<PEnt> h((size_t)EF * n, PEnt{0., 0, 0, 0, 0});
            for (int r = 1; r <= n; ++r)
                for (int e = t->h_row_ptr[r]; e < t->h_row_ptr[r + 1]; ++e) {
                    const Entry &en = t->h_ent[e];
                    PEnt &o = h[(size_t)(e - t->h_row_ptr[r]) * n + (r - 1)];
                    o.v = en.v;
                    o.a = (unsigned short)(8 * (en.jk & 0xffffu));
                    o.b = (unsigned short)(8 * (en.jk >> 16));
                    o.c = rank == 5 ? (unsigned short)(8 * (en.lm & 0xffffu)) : 0;
                    o.d = rank == 5 ? (unsigned short)(8 * (en.lm >> 16)) : 0;
                }
            // rank 3 leaves two index fields free: entry 0 of a row carries the row's true length, so that a thread
            // stops at the end of ITS row instead of walking the padding up to the longest one (MAOOAM-36: 351 entries
            // in 36 rows of up to 15)
            if (rank == 3)
                for (int r = 1; r <= n; ++r) h[r - 1].d = (unsigned short)(t->h_row_ptr[r + 1] - t->h_row_ptr[r]);
            pc->f.alloc(h.size());
            QGSB_CUDA(cudaMemcpy(pc->f.p, h.data(), h.size() * sizeof(PEnt), cudaMemcpyHostToDevice));
            pc->tab.f_ent = pc->f.p;
            pc->tab.EF = EF;
        }
    }
    // Jacobian positions, in the list order of the product policy (not needed by a value-baked bilinear product)
    if (!(spec && t->spec && t->spec->jac_hash != 0)) {
        const int npos = (int)t->h_pos_i.size();
        std::vector<int> order(npos);        // order[q] = position p served by list entry q
        std::vector<unsigned short> slots(npos);
        if (spec) {
            std::vector<std::pair<int, int>> by_slot(npos);
            for (int p = 0; p < npos; ++p)
                by_slot[p] = {t->spec->jac_slot_table[(t->h_pos_i[p] - 1) * n + (t->h_pos_j[p] - 1)], p};
            std::sort(by_slot.begin(), by_slot.end());
            for (int q = 0; q < npos; ++q) {
                order[q] = by_slot[q].second;
                slots[q] = (unsigned short)by_slot[q].first;
            }
        } else {
            for (int p = 0; p < npos; ++p) {
                order[p] = p;
                slots[p] = (unsigned short)((t->h_pos_i[p] - 1) * n + (t->h_pos_j[p] - 1));
            }
        }
        int EJ = 0;
        long real = 0;
        for (int p = 0; p < npos; ++p) {
            const int len = t->h_pos_ptr[p + 1] - t->h_pos_ptr[p];
            EJ = std::max(EJ, len);
            real += len;
        }
        if (ell_ok(EJ, npos, real)) {
            std::vector<PEnt> h((size_t)EJ * npos, PEnt{0., 0, 0, 0, 0});
            for (int q = 0; q < npos; ++q) {
                const int p = order[q];
                for (int e = t->h_pos_ptr[p]; e < t->h_pos_ptr[p + 1]; ++e) {
                    const Entry &en = t->h_jent[e];
                    PEnt &o = h[(size_t)(e - t->h_pos_ptr[p]) * npos + q];
                    o.v = en.v;
                    o.a = (unsigned short)(8 * (en.jk & 0xffffu));
                    o.b = rank == 5 ? (unsigned short)(8 * (en.jk >> 16)) : 0;
                    o.c = rank == 5 ? (unsigned short)(8 * (en.lm & 0xffffu)) : 0;
                }
                h[q].d = (unsigned short)(8 * slots[q]);      // entry 0 of the position carries its slot
            }
            pc->j.alloc(h.size());
            QGSB_CUDA(cudaMemcpy(pc->j.p, h.data(), h.size() * sizeof(PEnt), cudaMemcpyHostToDevice));
            pc->tab.j_ent = pc->j.p;
            pc->tab.EJ = EJ;
            pc->tab.npos = npos;
        }
    }
    slot = pc;
    return pc->tab;
}

void launch_pack_tangent(const qgsb_tensor *t, const TgParams &P, bool lyap)
{
    cudaError_t err;
    if (spec_tangent_usable(t) && forced_kernel() != 2) {
        err = t->spec->tangent(t->view, P, pack_tables(t, true), lyap ? 1 : 0, ctx().smem_optin, ctx().stream);
    } else {
        const PackTables &tab = pack_tables(t, false);
        switch (t->view.n) {
            case 20: err = launch_pack_dense_20(t->view, P, tab, lyap, ctx().smem_optin, ctx().stream); break;
            case 36: err = launch_pack_dense_36(t->view, P, tab, lyap, ctx().smem_optin, ctx().stream); break;
            case 38: err = launch_pack_dense_38(t->view, P, tab, lyap, ctx().smem_optin, ctx().stream); break;
            default: err = cudaErrorInvalidValue;
        }
    }
    count_launch();
    QGSB_REQUIRE(err == cudaSuccess, "packed tangent kernel launch failed: %s", cudaGetErrorString(err));
}

}  // namespace qgsb
