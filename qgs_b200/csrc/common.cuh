// common.cuh -- shared declarations of libqgsb (context, error convention, device tensor layout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qgsb.h"

namespace qgsb {

// ------------------------------------------------------------------------------------------------
// error convention: every extern "C" entry point returns 0 / non-zero and stores a message
// ------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
struct Failure {};  // thrown internally, caught at the ABI boundary

#define QGSB_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            qgsb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__,  \
                            __LINE__);                                                           \
            throw qgsb::Failure();                                                               \
        }                                                                                        \
    } while (0)

#define QGSB_REQUIRE(cond, ...)               \
    do {                                      \
        if (!(cond)) {                        \
            qgsb::set_error(__VA_ARGS__);     \
            throw qgsb::Failure();            \
        }                                     \
    } while (0)

// Entry points serialise on one process-wide lock, so the library may be called from several host threads (ctypes
// releases the GIL) -- calls queue up like the reference's integrate() calls queue on its worker pool.
std::recursive_mutex &api_mutex();
#define QGSB_API_LOCK std::lock_guard<std::recursive_mutex> api_guard__(qgsb::api_mutex());

#define QGSB_API_BEGIN \
    QGSB_API_LOCK      \
    try {
#define QGSB_API_END                                       \
    return 0;                                              \
    }                                                      \
    catch (const qgsb::Failure &) { return 1; }            \
    catch (const std::exception &e) {                      \
        qgsb::set_error("exception: %s", e.what());        \
        return 2;                                          \
    }

// ------------------------------------------------------------------------------------------------
// device contexts.  The library drives one OR SEVERAL devices from one process: slot 0 is the primary device (every
// handle is created there), further slots are the other devices of the box.  A context owns its streams, timing
// events and scratch pool.  ctx() is the context the CALLING THREAD is bound to (slot 0 unless the thread is one of
// the per-device workers of run_sharded): the single-device code below the entry points never needs to know that
// other devices exist.  This replaces the reference's pool of worker processes fed one trajectory at a time
// (qgs/integrators/integrator.py:121-142, 386-395).
// ------------------------------------------------------------------------------------------------
constexpr int MAX_DEVICES = 16;

struct PoolBlock {
    void *p;
    size_t bytes;
    bool used;
};

struct Context {
    bool ready = false;
    int slot = 0;
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    size_t smem_optin = 0;  // max dynamic shared memory per block
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // H2D / D2H streams of the pipelined host-buffer paths
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long launches = 0;
    std::vector<PoolBlock> pool;                          // grow-only scratch pool of this device
    // Pinned, device-mapped staging area for calls on a handful of states (f(t, x), Df(t, x) on one state): the kernel
    // reads its input from it and writes its result into it over the bus -- no cudaMemcpy in either direction.
    double *h_stage = nullptr;
    static constexpr size_t STAGE_IN = 4096, STAGE_OUT = 12288;   // doubles
};
Context &ctx();
void ensure_init();
inline void count_launch(long n = 1) { ctx().launches += n; }

// number of devices the library drives (>= 1 after ensure_init)
int device_slots();
// Contiguous balanced split of [0, n) over `parts`: part g is [g n / parts, (g + 1) n / parts)
// (the rule of qgs_b200.ensemble.shard_bounds and of SURVEY.md section 8e).
inline void shard_range(long n, int parts, int g, long *lo, long *hi)
{
    *lo = (long)((__int128)g * n / parts);
    *hi = (long)((__int128)(g + 1) * n / parts);
}
// How many devices a call over n_members should use: all of them when every shard keeps at least min_per_device
// members (so that the shards run the same kernel family as the whole ensemble would), else fewer, else 1.
int shard_count(long n_members, long min_per_device);
// Runs body(g, lo, hi) for the `parts` shards of [0, n) concurrently, shard g on device slot g: shard 0 on the calling
// thread, the others on worker threads bound to their device for the duration of the call (ctx() is theirs).  A
// failure in any shard is re-thrown on the calling thread with that shard's message.
void run_sharded(long n, int parts, const std::function<void(int, long, long)> &body);

// ------------------------------------------------------------------------------------------------
// ensemble layout in HBM ("tiled structure of arrays"): members are grouped in tiles of TILE = 128;
// inside a tile variable i of member m sits at i * TILE + (m % TILE), tiles follow each other with
// stride rows * TILE.  A thread block that owns a tile reads and writes one contiguous
// rows * 1 KiB chunk with fully coalesced 8-byte accesses, and every per-variable offset inside the
// tile is a compile-time constant for the tensor-specialised kernels.  ld = n_members rounded up to
// TILE; padding members are zero-filled.
// ------------------------------------------------------------------------------------------------
constexpr int TILE = 128;
__host__ __device__ inline size_t tile_base(long member, long rows)
{
    return (size_t)(member / TILE) * (size_t)rows * TILE + (size_t)(member % TILE);
}

// RAII device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) QGSB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void upload(const T *h, size_t count, cudaStream_t s) {
        QGSB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count, cudaStream_t s) const {
        QGSB_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// Scratch buffer drawn from a grow-only pool owned by the context: the host-buffer entry points
// run many times with the same sizes (one integrate() per chunk of a long run), and cudaMalloc /
// cudaFree of hundreds of MB per call would dominate their end-to-end time.
void *pool_acquire(size_t bytes);
void pool_release(Context *owner, void *p);
void pool_trim();
template <typename T>
struct PoolBuf {
    T *p = nullptr;
    size_t n = 0;
    Context *owner = nullptr;   // the device context whose pool the buffer came from (and goes back to)
    PoolBuf() {}
    explicit PoolBuf(size_t count) { alloc(count); }
    PoolBuf(const PoolBuf &) = delete;
    PoolBuf &operator=(const PoolBuf &) = delete;
    ~PoolBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        owner = &ctx();
        p = (T *)pool_acquire(std::max<size_t>(count, 1) * sizeof(T));
    }
    void release() {
        if (p) pool_release(owner, p);
        p = nullptr;
        n = 0;
    }
    void upload(const T *h, size_t count, cudaStream_t s) {
        QGSB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count, cudaStream_t s) const {
        QGSB_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// ------------------------------------------------------------------------------------------------
// device tensor layout
//
// One 16-byte entry per non-zero, rows (first index) sorted ascending with CSR row pointers so that
// every thread walks the same entry stream: {value, j | k << 16, l | m << 16}.  Rank-3 entries use
// l = m = 0, i.e. they multiply by x_0 = 1 like the reference's rank-5 tensors do for low orders.
// ------------------------------------------------------------------------------------------------
struct __align__(16) Entry {
    double v;
    uint32_t jk;
    uint32_t lm;
};

// Jacobian in "position" form: the distinct (i, j >= 1) matrix positions, CSR by row i (for J @ X)
// and CSC by column j (for J^T @ X); every position owns a contiguous run of entries
// {value, k | l << 16, m} whose products are summed to give J_ij.
struct JacView {
    int npos = 0;               // number of structurally non-zero J_ij, i,j in 1..n
    const int *pos_ptr = nullptr;    // (npos + 1) run boundaries into ent
    const Entry *ent = nullptr;      // jk = k | l << 16, lm = m
    const int *pos_i = nullptr;      // (npos) row index   (1-based variable index)
    const int *pos_j = nullptr;      // (npos) column index
    const int *row_ptr = nullptr;    // (n + 2) CSR over positions, by i
    const int *col_ptr = nullptr;    // (n + 2) CSC over positions, by j
    const int *col_perm = nullptr;   // (npos) position ids ordered by (j, i)
};

struct TensorView {
    int n = 0;        // ndim
    int rank = 3;
    int nnz = 0;
    const Entry *ent = nullptr;   // (nnz) sorted by row
    const int *row_ptr = nullptr; // (n + 2): entries of row i are [row_ptr[i], row_ptr[i+1])
    JacView jac;
};

struct SpecKernels;  // tensor-specialised kernels (spec_registry.h)

}  // namespace qgsb

struct qgsb_tensor {
    qgsb::TensorView view;
    int max_deg = 2;           // largest number of non-trivial factors of an entry
    int jac_max_deg = 1;
    long nnz_in = 0, jnnz_in = 0;
    uint64_t hash = 0;
    std::vector<double> val_sorted;      // values in device entry order (for the specialised kernels)
    std::vector<int32_t> coo_sorted;     // (nnz, rank) in device entry order
    qgsb::DevBuf<qgsb::Entry> d_ent, d_jent;
    qgsb::DevBuf<int> d_row_ptr, d_pos_ptr, d_pos_i, d_pos_j, d_jrow_ptr, d_jcol_ptr, d_jcol_perm;
    const qgsb::SpecKernels *spec = nullptr;
    bool use_spec = true;
    bool jac_matches_spec = false;       // every Jacobian position of this handle has a slot in spec->jac_slot_table
    std::vector<int> h_pos_i, h_pos_j;   // host copy of the Jacobian positions
    std::vector<qgsb::Entry> h_ent, h_jent;          // host copies of the device entry lists
    std::vector<int> h_row_ptr, h_pos_ptr;
    mutable qgsb::DevBuf<int> d_wide;    // row-to-warp assignment of the wide RK kernel (rk.cu), built on first use
    struct G3Cache;                      // plan + tables of the large-basis RK kernel (rk.cu G3), built on first use
    mutable G3Cache *g3_cache = nullptr;
    mutable int g3_state = 0;            // 0 not planned, 1 ready, -1 not applicable
    // tables of the packed tangent kernels, built on first use: [0] dense product, [1] generated product
    struct PackCache;
    mutable PackCache *pack_cache[2] = {nullptr, nullptr};
    // ---- several devices: the handle lives on `device`; the other devices get a replica on first use ----
    int device = 0;                                  // CUDA ordinal the device arrays above live on
    int ndim_in = 0;                                 // the caller's arrays, kept to build replicas
    std::vector<int32_t> in_coo, in_jcoo;
    std::vector<double> in_val, in_jval;
    mutable std::map<int, qgsb_tensor *> replicas;   // CUDA ordinal -> copy of this handle on that device
    mutable std::mutex replica_mutex;
    ~qgsb_tensor();
};

namespace qgsb {
void g3_release(qgsb_tensor::G3Cache *c);
// the copy of handle t that lives on the calling thread's device (t itself on its own device)
const qgsb_tensor *tensor_here(const qgsb_tensor *t);
}

struct qgsb_ensemble {
    const qgsb_tensor *tensor = nullptr;
    long N = 0, ld = 0;
    int device = 0;                  // CUDA ordinal the arrays below live on
    qgsb::DevBuf<double> d_y;        // (n, ld)
    qgsb::DevBuf<double> d_stage;    // scratch for AoS <-> SoA staging (N * n)
    qgsb::DevBuf<double> d_dt;       // step lengths of the launch in flight
    // A large ensemble created while the library drives several devices is a COMPOSITE: its members live in contiguous
    // blocks [lo[g], lo[g + 1]) on the devices, each block an ordinary single-device ensemble; the handle itself then
    // owns no device memory and every call fans out over the parts (run_sharded).
    std::vector<qgsb_ensemble *> parts;
    std::vector<long> lo;
    ~qgsb_ensemble();
};
