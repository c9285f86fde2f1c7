// tgls_pack.cu -- dispatch of the packed tangent-linear / Benettin kernels (tgls_pack.cuh).
//
// A tensor with a generated module (spec_registry.h: SpecKernels::tangent) runs the product over its
// literal Jacobian position list; every other tensor of a supported ndim runs the dense n x n product.
#include <cstdlib>
#include <cstring>

#include "spec_registry.h"
#include "tgls_pack.cuh"

namespace qgsb {

static int forced_kernel()
{
    // QGSB_TGLS_KERNEL = pack | pack_dense | reg | generic  (testing / A-B measurements)
    const char *e = getenv("QGSB_TGLS_KERNEL");
    if (!e || !e[0]) return 0;
    if (!strcmp(e, "pack")) return 1;
    if (!strcmp(e, "pack_dense")) return 2;
    if (!strcmp(e, "reg")) return 3;
    if (!strcmp(e, "generic")) return 4;
    return 0;
}

template <int N>
static cudaError_t launch_dense(const TensorView &T, const TgParams &P, bool lyap)
{
    return pack::launch<N, pack::DenseProduct<N, false>, pack::DenseProduct<N, true>>(T, P, lyap, ctx().smem_optin,
                                                                                       ctx().stream);
}

static bool dense_ndim(int n) { return n == 20 || n == 36 || n == 38; }

static bool spec_tangent_usable(const qgsb_tensor *t)
{
    return t->spec && t->use_spec && t->spec->tangent && t->jac_matches_spec;
}

bool pack_tangent_supported(const qgsb_tensor *t, const Tableau &tab, int m)
{
    const int force = forced_kernel();
    if (force == 3 || force == 4) return false;
    if (!tab.chain || m < 1 || m > pack::MAX_THREADS) return false;
    const int n = t->view.n;
    if (!(spec_tangent_usable(t) && force != 2) && !dense_ndim(n)) return false;
    // a block must hold enough columns to fill its warps: with very few columns per member the per-member shared
    // state limits the packing and the generic kernel (one block per member) is the better fit
    const int jv = (spec_tangent_usable(t) && force != 2) ? t->spec->jac_slots : n * n;
    const size_t per_member = (size_t)(jv + 8 * n + 2 * (size_t)n * m + 3 * m + 32) * sizeof(double);
    const int G = std::min<int>(pack::MAX_THREADS / m, (int)(ctx().smem_optin / per_member));
    return G >= 1 && G * m >= 96;
}

void launch_pack_tangent(const qgsb_tensor *t, const TgParams &P, bool lyap)
{
    cudaError_t err;
    if (spec_tangent_usable(t) && forced_kernel() != 2) {
        err = t->spec->tangent(t->view, P, lyap ? 1 : 0, ctx().smem_optin, ctx().stream);
    } else {
        switch (t->view.n) {
            case 20: err = launch_dense<20>(t->view, P, lyap); break;
            case 36: err = launch_dense<36>(t->view, P, lyap); break;
            case 38: err = launch_dense<38>(t->view, P, lyap); break;
            default: err = cudaErrorInvalidValue;
        }
    }
    count_launch();
    QGSB_REQUIRE(err == cudaSuccess, "packed tangent kernel launch failed: %s", cudaGetErrorString(err));
}

}  // namespace qgsb
