// tgls_pack_dense_36.cu -- the packed tangent-linear / Benettin kernels with the dense n x n product for ndim = 36
// (any tensor of that size without a generated module).  One translation unit per ndim: they compile in parallel.
#include "tgls_pack.cuh"

namespace qgsb {

cudaError_t launch_pack_dense_36(const TensorView &T, const TgParams &P, const PackTables &tab, bool lyap, size_t smem,
                                 cudaStream_t stream)
{
    return pack::launch<36, pack::DenseProduct<36, false>, pack::DenseProduct<36, true>>(T, P, tab, lyap, smem, stream);
}

}  // namespace qgsb
