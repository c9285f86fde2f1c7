// tgls_pack_dense_20.cu -- the packed tangent-linear / Benettin kernels with the dense n x n product for ndim = 20
// (any tensor of that size without a generated module).  One translation unit per ndim: they compile in parallel.
#include "tgls_pack.cuh"

namespace qgsb {

cudaError_t launch_pack_dense_20(const TensorView &T, const TgParams &P, const PackTables &tab, bool lyap, size_t smem,
                                 cudaStream_t stream)
{
    return pack::launch<20, pack::DenseProduct<20, false>, pack::DenseProduct<20, true>>(T, P, tab, lyap, smem, stream);
}

}  // namespace qgsb
