// spec_registry.h -- interface between libqgsb and tensor-specialised kernel modules.
//
// qgs_b200/codegen.py turns one tensor (ndim, rank, row-sorted index list AND values) into
// straight-line sm_100a code whose state lives in registers: indices are compile-time constants and
// the tensor values are immediates of the instruction stream.  A module registers itself either from
// a static initialiser when it is linked into libqgsb.so (the canonical configurations, built by
// __graft_entry__.build()) or through its exported qgsb_plugin_kernels() when it is a separate
// shared object loaded with qgsb_load_plugin() (run-time builds for other parameter sets).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qgsb {

struct TensorView;   // common.cuh
struct TgParams;     // tgls_shared.cuh
struct PackTables;   // tgls_shared.cuh

// Layout version of SpecKernels / TgParams / PackTables / the packed kernels' shared-memory carve-up as a module was
// compiled against them.  Bump it whenever one of those changes: qgsb_load_plugin refuses a module built against
// another version instead of reading its table with the wrong layout (modules cached on disk outlive a library).
#define QGSB_SPEC_ABI 8u

struct SpecKernels {
    uint32_t abi;   // QGSB_SPEC_ABI of the headers the module was compiled with -- must stay the FIRST member
    uint64_t hash;  // FNV-1a over (ndim, rank, nnz, row-sorted coo, values) -- see tensor_hash()
    int n, rank, nnz;
    const char *name;
    // "chain" Runge-Kutta (a_ij != 0 only for j = i-1: Euler, midpoint, Heun, classic RK4 ...):
    // alpha[i] = a[i][i-1] (alpha[0] unused), beta[i] = b[i].  d_y is the tiled-SoA state.
    // d_rec == nullptr: no recording; else records laid out (R, tiled SoA) as qgsb_rk_integrate defines.
    cudaError_t (*rk_chain)(double *d_y, long ld, long n_members, long n_steps, const double *d_dt,
                            int s, const double *alpha, const double *beta, long write_steps,
                            long n_records, double *d_rec, int sm_count, cudaStream_t stream);
    // batched tendencies on tiled-SoA arrays: out = f(x)
    cudaError_t (*tendencies)(const double *d_x, double *d_out, long ld, long n_members,
                              cudaStream_t stream);
    // packed tangent-linear / Benettin kernels (tgls_pack.cuh) with the product J @ X emitted over the literal
    // list of Jacobian positions; null when the module was generated without a Jacobian tensor.
    // lyap = 0: integrate.py:555-614, 1: lyapunov.py:471-632.
    cudaError_t (*tangent)(const TensorView &T, const TgParams &P, const PackTables &tables, int lyap,
                           size_t smem_limit, cudaStream_t stream);
    int jac_slots;                 // doubles of shared memory holding the position values of one member
    const short *jac_slot_table;   // (n, n) row-major: slot of position (i, j), -1 where J_ij is structurally zero
    uint64_t jac_hash;             // != 0: the tangent product has the Jacobian tensor's VALUES baked in (bilinear
                                   // form); FNV-1a over its (i, j)-sorted entries -- see jacobian_hash()
    // Runge-Kutta with ANY explicit tableau (a (s, s) row-major, j < i read; integrate.py:214-219): the stage
    // derivatives live in shared memory.  Null in modules too large to carry a second copy of f; returns
    // cudaErrorInvalidValue when s + 1 state sets exceed smem_limit (the caller then uses the generic kernel).
    cudaError_t (*rk_general)(double *d_y, long ld, long n_members, long n_steps, const double *d_dt, int s,
                              const double *a, const double *beta, long write_steps, long n_records, double *d_rec,
                              size_t smem_limit, cudaStream_t stream);
};

void register_spec(const SpecKernels *k);
const SpecKernels *find_spec(uint64_t hash);
uint64_t tensor_hash(int n, int rank, long nnz, const int32_t *coo_sorted, const double *val_sorted);

struct SpecRegistrar {
    explicit SpecRegistrar(const SpecKernels *k) { register_spec(k); }
};

}  // namespace qgsb

// a plugin exports:   extern "C" const qgsb::SpecKernels *qgsb_plugin_kernels(void);
