// runtime.cu -- context, error convention, tensor hand-off, layout kernels, batched f / Df and the
// raw sparse_mul* contractions of libqgsb.
#include <dlfcn.h>
#include <stdarg.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <thread>

#include "common.cuh"
#include "kernels.cuh"
#include "spec_registry.h"

namespace qgsb {

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// device contexts
// ------------------------------------------------------------------------------------------------
static Context g_ctx[MAX_DEVICES];
static int g_slots = 0;                      // devices the library drives; 0 = not initialised
static thread_local Context *tl_ctx = nullptr;

Context &ctx() { return tl_ctx ? *tl_ctx : g_ctx[0]; }
int device_slots() { return g_slots; }

std::recursive_mutex &api_mutex()
{
    static std::recursive_mutex m;
    return m;
}

static void context_open(Context &c, int slot, int device)
{
    QGSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    QGSB_CUDA(cudaGetDeviceProperties(&p, device));
    c.slot = slot;
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    c.cc_major = p.major;
    c.cc_minor = p.minor;
    c.total_mem = p.totalGlobalMem;
    c.smem_optin = p.sharedMemPerBlockOptin;
    QGSB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    QGSB_CUDA(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking));
    QGSB_CUDA(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking));
    QGSB_CUDA(cudaEventCreate(&c.ev0));
    QGSB_CUDA(cudaEventCreate(&c.ev1));
    if (cudaHostAlloc((void **)&c.h_stage, (Context::STAGE_IN + Context::STAGE_OUT) * sizeof(double),
                      cudaHostAllocMapped) != cudaSuccess) {
        c.h_stage = nullptr;            // the single-state calls then take the ordinary copies
        cudaGetLastError();
    }
    c.ready = true;
}

static void context_close(Context &c)
{
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    for (auto &blk : c.pool)
        if (!blk.used) cudaFree(blk.p);
    c.pool.erase(std::remove_if(c.pool.begin(), c.pool.end(), [](const PoolBlock &b) { return !b.used; }), c.pool.end());
    cudaStreamDestroy(c.own_stream);
    cudaStreamDestroy(c.copy_in);
    cudaStreamDestroy(c.copy_out);
    cudaEventDestroy(c.ev0);
    cudaEventDestroy(c.ev1);
    if (c.h_stage) cudaFreeHost(c.h_stage);
    c.h_stage = nullptr;
    c.own_stream = c.stream = c.copy_in = c.copy_out = nullptr;
    c.ev0 = c.ev1 = nullptr;
    c.ready = false;
}

void raw_cache_clear();

static void close_all()
{
    raw_cache_clear();
    for (int g = 0; g < MAX_DEVICES; ++g) context_close(g_ctx[g]);
    g_slots = 0;
}

// Which devices does the process drive?  An explicit ordinal: that one.  Otherwise QGSB_DEVICES ("all", a count, or a
// comma-separated list of ordinals) decides; without it a process started by torchrun (LOCAL_RANK set: one process
// per GPU) takes its own device, and any other process takes EVERY visible device -- one integrate() call then
// spreads its members over the box like the reference's spreads them over its worker processes.
static std::vector<int> choose_devices(int device, int count)
{
    std::vector<int> devs;
    if (device >= 0) {
        QGSB_REQUIRE(device < count, "device %d requested but only %d visible", device, count);
        devs.push_back(device);
        return devs;
    }
    const char *env = getenv("QGSB_DEVICES");
    const char *lr = getenv("LOCAL_RANK");
    if (env && env[0]) {
        if (!strcmp(env, "all")) {
            for (int d = 0; d < count; ++d) devs.push_back(d);
        } else if (strchr(env, ',')) {
            std::string list(env);
            size_t pos = 0;
            while (pos <= list.size()) {
                size_t next = list.find(',', pos);
                if (next == std::string::npos) next = list.size();
                if (next > pos) devs.push_back(atoi(list.substr(pos, next - pos).c_str()));
                pos = next + 1;
            }
        } else {
            const int want = atoi(env);
            // a single number is a COUNT of devices ("QGSB_DEVICES=1": stay on one device)
            for (int d = 0; d < std::min(std::max(want, 1), count); ++d) devs.push_back(d);
        }
    } else if (lr) {
        devs.push_back(atoi(lr) % count);
    } else {
        for (int d = 0; d < count; ++d) devs.push_back(d);
    }
    QGSB_REQUIRE(!devs.empty() && (int)devs.size() <= MAX_DEVICES, "QGSB_DEVICES selects %zu devices (1..%d allowed)",
                 devs.size(), MAX_DEVICES);
    for (size_t q = 0; q < devs.size(); ++q) {
        QGSB_REQUIRE(devs[q] >= 0 && devs[q] < count, "device %d requested but only %d visible", devs[q], count);
        for (size_t r = 0; r < q; ++r) QGSB_REQUIRE(devs[r] != devs[q], "device %d listed twice", devs[q]);
    }
    return devs;
}

static void init_devices(const std::vector<int> &devs)
{
    bool same = g_slots == (int)devs.size();
    for (int g = 0; same && g < g_slots; ++g) same = g_ctx[g].device == devs[g];
    if (same && g_ctx[0].ready) {
        cudaSetDevice(g_ctx[0].device);
        return;
    }
    close_all();
    // the primary context now; the others when a sharded call first needs them (run_sharded)
    context_open(g_ctx[0], 0, devs[0]);
    for (size_t g = 1; g < devs.size(); ++g) {
        g_ctx[g].slot = (int)g;
        g_ctx[g].device = devs[g];
    }
    g_slots = (int)devs.size();
}

static void init_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    QGSB_REQUIRE(e == cudaSuccess && count > 0,
                 "no CUDA device available (%s); libqgsb has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    init_devices(choose_devices(device, count));
}

void ensure_init()
{
    if (g_slots == 0 || !g_ctx[0].ready) init_device(-1);
    else cudaSetDevice(ctx().device);
}

int shard_count(long n_members, long min_per_device)
{
    if (g_slots <= 1 || tl_ctx != nullptr) return 1;      // a worker never shards again
    const long by_size = n_members / std::max<long>(min_per_device, 1);
    return (int)std::max<long>(1, std::min<long>(g_slots, by_size));
}

void run_sharded(long n, int parts, const std::function<void(int, long, long)> &body)
{
    if (parts <= 1) {
        body(0, 0, n);
        return;
    }
    std::vector<std::string> errors(parts);
    std::vector<int> failed(parts, 0);
    auto shard = [&](int g) {
        long lo, hi;
        shard_range(n, parts, g, &lo, &hi);
        try {
            Context &c = g_ctx[g];
            if (!c.ready) context_open(c, g, c.device);
            else QGSB_CUDA(cudaSetDevice(c.device));
            tl_ctx = &c;
            if (hi > lo) body(g, lo, hi);
        } catch (const Failure &) {
            failed[g] = 1;
            errors[g] = g_err;
        } catch (const std::exception &e) {
            failed[g] = 1;
            errors[g] = std::string("exception: ") + e.what();
        }
        tl_ctx = nullptr;
    };
    std::vector<std::thread> workers;
    for (int g = 1; g < parts; ++g) workers.emplace_back(shard, g);
    shard(0);
    for (auto &w : workers) w.join();
    cudaSetDevice(g_ctx[0].device);
    for (int g = 0; g < parts; ++g)
        if (failed[g]) {
            set_error("device %d (shard %d of %d): %s", g_ctx[g].device, g, parts, errors[g].c_str());
            throw Failure();
        }
}

// ------------------------------------------------------------------------------------------------
// scratch pool (one per device context)
// ------------------------------------------------------------------------------------------------
void pool_trim()
{
    auto &b = ctx().pool;
    for (size_t i = 0; i < b.size();) {
        if (!b[i].used) {
            cudaFree(b[i].p);
            b.erase(b.begin() + i);
        } else {
            ++i;
        }
    }
}

void *pool_acquire(size_t bytes)
{
    auto &b = ctx().pool;
    int best = -1;
    for (size_t i = 0; i < b.size(); ++i)
        if (!b[i].used && b[i].bytes >= bytes && (best < 0 || b[i].bytes < b[best].bytes)) best = (int)i;
    if (best >= 0 && b[best].bytes <= 2 * bytes + (1 << 20)) {
        b[best].used = true;
        return b[best].p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim();  // give cached blocks back and retry once
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        throw Failure();
    }
    b.push_back({p, bytes, true});
    return p;
}

void pool_release(Context *owner, void *p)
{
    // a buffer goes back to the pool of the device it came from; only the thread bound to that device touches it
    for (auto &blk : owner->pool)
        if (blk.p == p) {
            blk.used = false;
            return;
        }
}

// ------------------------------------------------------------------------------------------------
// specialised-kernel registry
// ------------------------------------------------------------------------------------------------
static std::map<uint64_t, SpecKernels> &registry()
{
    static std::map<uint64_t, SpecKernels> r;
    return r;
}

// A module may carry the Runge-Kutta part, the tangent part or both; parts registered for the same tensor are merged.
void register_spec(const SpecKernels *k)
{
    auto it = registry().find(k->hash);
    if (it == registry().end()) {
        registry()[k->hash] = *k;
        return;
    }
    SpecKernels &m = it->second;
    if (k->rk_chain) m.rk_chain = k->rk_chain;
    if (k->rk_general) m.rk_general = k->rk_general;
    if (k->tendencies) m.tendencies = k->tendencies;
    if (k->tangent) {
        m.tangent = k->tangent;
        m.jac_slots = k->jac_slots;
        m.jac_slot_table = k->jac_slot_table;
        m.jac_hash = k->jac_hash;
    }
}

const SpecKernels *find_spec(uint64_t hash)
{
    auto it = registry().find(hash);
    return it == registry().end() ? nullptr : &it->second;
}

// FNV-1a over the (i, j)-sorted Jacobian entries kept by prepare_mat: all index tuples (int32), then all values
static uint64_t jacobian_hash(const qgsb_tensor *t)
{
    uint64_t h = 1469598103934665603ULL;
    auto mix = [&h](const void *p, size_t bytes) {
        const unsigned char *c = (const unsigned char *)p;
        for (size_t q = 0; q < bytes; ++q) {
            h ^= (uint64_t)c[q];
            h *= 1099511628211ULL;
        }
    };
    const int rank = t->view.rank;
    for (size_t p = 0; p < t->h_pos_i.size(); ++p)
        for (int e = t->h_pos_ptr[p]; e < t->h_pos_ptr[p + 1]; ++e) {
            const Entry &en = t->h_jent[e];
            int32_t idx[5] = {t->h_pos_i[p], t->h_pos_j[p], 0, 0, 0};
            if (rank == 5) {
                idx[2] = (int32_t)(en.jk & 0xffffu);
                idx[3] = (int32_t)(en.jk >> 16);
                idx[4] = (int32_t)en.lm;
            } else {
                idx[2] = (int32_t)en.jk;
            }
            mix(idx, sizeof(int32_t) * rank);
        }
    for (const Entry &en : t->h_jent) mix(&en.v, sizeof(double));
    return h;
}

static bool jacobian_matches(const qgsb_tensor *t)
{
    const SpecKernels *k = t->spec;
    if (!k || !k->tangent || !k->jac_slot_table) return false;
    if (k->jac_hash != 0) return jacobian_hash(t) == k->jac_hash;      // value-baked product: exact tensor only
    const int n = t->view.n;
    for (size_t p = 0; p < t->h_pos_i.size(); ++p) {
        const int i = t->h_pos_i[p], j = t->h_pos_j[p];
        if (i < 1 || j < 1 || i > n || j > n || k->jac_slot_table[(i - 1) * n + (j - 1)] < 0) return false;
    }
    return true;
}

uint64_t tensor_hash(int n, int rank, long nnz, const int32_t *coo_sorted, const double *val_sorted)
{
    uint64_t h = 1469598103934665603ULL;
    auto mix_bytes = [&h](const void *p, size_t bytes) {
        const unsigned char *c = (const unsigned char *)p;
        for (size_t q = 0; q < bytes; ++q) {
            h ^= (uint64_t)c[q];
            h *= 1099511628211ULL;
        }
    };
    const int32_t head[3] = {n, rank, (int32_t)nnz};
    mix_bytes(head, sizeof(head));
    mix_bytes(coo_sorted, sizeof(int32_t) * (size_t)nnz * rank);
    mix_bytes(val_sorted, sizeof(double) * (size_t)nnz);
    return h;
}

// ------------------------------------------------------------------------------------------------
// host-side tensor preparation
// ------------------------------------------------------------------------------------------------
struct HostTensor {
    std::vector<Entry> ent;
    std::vector<int> row_ptr;
    std::vector<int32_t> coo_sorted;
    std::vector<double> val_sorted;
};

static void check_indices(int n1, int rank, long nnz, const int32_t *coo)
{
    QGSB_REQUIRE(rank == 3 || rank == 5, "tensor rank must be 3 or 5, got %d", rank);
    QGSB_REQUIRE(n1 >= 1 && n1 <= 65535, "dimension %d out of range (1..65535)", n1);
    QGSB_REQUIRE(nnz >= 0 && nnz < (1L << 31) / 8, "nnz %ld out of range", nnz);
    for (long e = 0; e < nnz * rank; ++e)
        QGSB_REQUIRE(coo[e] >= 0 && coo[e] < n1, "tensor index %d outside [0, %d) at entry %ld", coo[e], n1,
                     e / rank);
}

// tendencies form: entries stably sorted by row, CSR over rows 0..n1-1
static HostTensor prepare_vec(int n1, int rank, long nnz, const int32_t *coo, const double *val)
{
    check_indices(n1, rank, nnz, coo);
    std::vector<long> order(nnz);
    std::iota(order.begin(), order.end(), 0L);
    std::stable_sort(order.begin(), order.end(),
                     [&](long x, long y) { return coo[x * rank] < coo[y * rank]; });
    HostTensor h;
    h.ent.resize(nnz);
    h.row_ptr.assign(n1 + 1, 0);
    h.coo_sorted.resize(nnz * rank);
    h.val_sorted.resize(nnz);
    for (long q = 0; q < nnz; ++q) {
        const int32_t *c = coo + order[q] * rank;
        Entry &e = h.ent[q];
        e.v = val[order[q]];
        e.jk = (uint32_t)c[1] | ((uint32_t)c[2] << 16);
        e.lm = rank == 5 ? ((uint32_t)c[3] | ((uint32_t)c[4] << 16)) : 0u;
        h.row_ptr[c[0] + 1]++;
        for (int r = 0; r < rank; ++r) h.coo_sorted[q * rank + r] = c[r];
        h.val_sorted[q] = e.v;
    }
    for (int i = 0; i < n1; ++i) h.row_ptr[i + 1] += h.row_ptr[i];
    return h;
}

struct HostJac {
    std::vector<Entry> ent;
    std::vector<int> pos_ptr, pos_i, pos_j, row_ptr, col_ptr, col_perm;
};

// matrix form: distinct (i, j) positions with their entry runs; keep_zero keeps row/column 0
static HostJac prepare_mat(int n1, int rank, long nnz, const int32_t *coo, const double *val, bool keep_zero)
{
    check_indices(n1, rank, nnz, coo);
    std::vector<long> order;
    order.reserve(nnz);
    for (long e = 0; e < nnz; ++e)
        if (keep_zero || (coo[e * rank] > 0 && coo[e * rank + 1] > 0)) order.push_back(e);
    std::stable_sort(order.begin(), order.end(), [&](long x, long y) {
        const int32_t *a = coo + x * rank, *b = coo + y * rank;
        return a[0] != b[0] ? a[0] < b[0] : a[1] < b[1];
    });
    HostJac h;
    h.ent.resize(order.size());
    h.row_ptr.assign(n1 + 1, 0);
    h.col_ptr.assign(n1 + 1, 0);
    int li = -1, lj = -1;
    for (size_t q = 0; q < order.size(); ++q) {
        const int32_t *c = coo + order[q] * rank;
        if (c[0] != li || c[1] != lj) {
            h.pos_ptr.push_back((int)q);
            h.pos_i.push_back(c[0]);
            h.pos_j.push_back(c[1]);
            h.row_ptr[c[0] + 1]++;
            h.col_ptr[c[1] + 1]++;
            li = c[0];
            lj = c[1];
        }
        Entry &e = h.ent[q];
        e.v = val[order[q]];
        e.jk = rank == 5 ? ((uint32_t)c[2] | ((uint32_t)c[3] << 16)) : (uint32_t)c[2];
        e.lm = rank == 5 ? (uint32_t)c[4] : 0u;
    }
    h.pos_ptr.push_back((int)order.size());
    for (int i = 0; i < n1; ++i) {
        h.row_ptr[i + 1] += h.row_ptr[i];
        h.col_ptr[i + 1] += h.col_ptr[i];
    }
    const int npos = (int)h.pos_i.size();
    h.col_perm.resize(npos);
    std::iota(h.col_perm.begin(), h.col_perm.end(), 0);
    std::stable_sort(h.col_perm.begin(), h.col_perm.end(),
                     [&](int x, int y) { return h.pos_j[x] < h.pos_j[y]; });
    return h;
}

template <typename T>
static void to_device(DevBuf<T> &d, const std::vector<T> &h)
{
    d.alloc(std::max<size_t>(h.size(), 1));
    if (!h.empty()) QGSB_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
}

// ------------------------------------------------------------------------------------------------
// kernels: layout changes
// ------------------------------------------------------------------------------------------------
// (N, n) member-major  ->  tiled SoA over ld members; padding members are zero-filled
__global__ void aos_to_soa_kernel(const double *__restrict__ in, double *__restrict__ out, long N, int n, long ld)
{
    __shared__ double tile[32][33];
    const long m0 = (long)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        long m = m0 + r;
        int i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (m < N && i < n) ? in[m * n + i] : 0.;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r;
        long m = m0 + threadIdx.x;
        if (i < n && m < ld) out[tile_base(m, n) + (size_t)i * TILE] = tile[threadIdx.x][r];
    }
}

__global__ void soa_to_aos_kernel(const double *__restrict__ in, double *__restrict__ out, long N, int n, long ld)
{
    __shared__ double tile[32][33];
    const long m0 = (long)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r;
        long m = m0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && m < N) ? in[tile_base(m, n) + (size_t)i * TILE] : 0.;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        long m = m0 + r;
        int i = i0 + threadIdx.x;
        if (m < N && i < n) out[m * n + i] = tile[threadIdx.x][r];
    }
}

// records (R, tiled SoA of `rows` variables) -> API layout (N, rows, R); flip reverses the record axis
// (integrate.py:223).  One (record, member) tile per block and per row.
__global__ void rec_to_api_kernel(const double *__restrict__ in, double *__restrict__ out, long N, int rows,
                                  long R, long ld, int flip, long m_tiles, int row0)
{
    __shared__ double tile[32][33];
    const long m0 = ((long)blockIdx.x % m_tiles) * 32;
    const long r0 = ((long)blockIdx.x / m_tiles) * 32;
    const int i = blockIdx.y + row0;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        long r = r0 + q;
        long m = m0 + threadIdx.x;
        tile[q][threadIdx.x] = (r < R && m < N) ? in[(size_t)r * rows * ld + tile_base(m, rows) + (size_t)i * TILE] : 0.;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        long m = m0 + q;
        long r = r0 + threadIdx.x;
        if (m < N && r < R) {
            long ro = flip ? R - 1 - r : r;
            out[(m * rows + i) * R + ro] = tile[threadIdx.x][q];
        }
    }
}

void launch_aos_to_soa(const double *d_in, double *d_out, long N, int n, long ld)
{
    dim3 grid((unsigned)((ld + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
    aos_to_soa_kernel<<<grid, block, 0, ctx().stream>>>(d_in, d_out, N, n, ld);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

void launch_soa_to_aos(const double *d_in, double *d_out, long N, int n, long ld)
{
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
    soa_to_aos_kernel<<<grid, block, 0, ctx().stream>>>(d_in, d_out, N, n, ld);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
}

void launch_rec_to_api(const double *d_in, double *d_out, long N, long rows, long R, long ld, int flip)
{
    const long m_tiles = (N + 31) / 32, r_tiles = (R + 31) / 32;
    QGSB_REQUIRE(m_tiles * r_tiles < (1L << 31), "record buffer too large for one layout-change launch");
    // rows can exceed the grid y limit (65535) for n * m fundamental matrices: chunk it
    for (long z0 = 0; z0 < rows; z0 += 65535) {
        long nz = std::min<long>(65535, rows - z0);
        dim3 grid((unsigned)(m_tiles * r_tiles), (unsigned)nz), block(32, 8);
        rec_to_api_kernel<<<grid, block, 0, ctx().stream>>>(d_in, d_out, N, (int)rows, R, ld, flip, m_tiles, (int)z0);
        count_launch();
    }
    QGSB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// kernels: batched contractions on member-major input (the f / Df closures and sparse_mul*)
// ------------------------------------------------------------------------------------------------
// one thread per (member, row): out[m][i - shift] = sum over row i.  The four vectors may differ
// (sparse_mul3/5 proper) or all alias x (the f closure).  stride = distance between members.
template <int RANK>
__global__ void mulvec_kernel(TensorView T, long N, int n1, const double *__restrict__ va,
                              const double *__restrict__ vb, const double *__restrict__ vc,
                              const double *__restrict__ vd, long stride, int with_x0, double *__restrict__ out,
                              int shift, int out_stride)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = n1 - shift;
    if (t >= N * rows) return;
    const long m = t / rows;
    const int i = (int)(t % rows) + shift;
    // with_x0: vectors hold only x_1..x_n and x_0 = 1 is implied (the closures' concatenate, tendencies.py:112)
    auto ld = [&](const double *v, uint32_t idx) -> double {
        if (with_x0) return idx == 0 ? 1. : v[m * stride + idx - 1];
        return v[m * stride + idx];
    };
    double acc = 0.;
    for (int e = T.row_ptr[i]; e < T.row_ptr[i + 1]; ++e) {
        const Entry en = T.ent[e];
        double p = ld(va, en.jk & 0xffffu) * ld(vb, en.jk >> 16);
        if (RANK == 5) p = p * ld(vc, en.lm & 0xffffu) * ld(vd, en.lm >> 16);
        acc += p * en.v;
    }
    if (i == 0) acc = 1.;  // sparse_mul.py:80 / :157
    out[m * out_stride + (i - shift)] = acc;
}

// one thread per (member, matrix position): out[m][i - shift][j - shift] = J_ij
template <int RANK>
__global__ void mulmat_kernel(JacView J, long N, const double *__restrict__ va, const double *__restrict__ vb,
                              const double *__restrict__ vc, long stride, int with_x0, double *__restrict__ out,
                              int shift, int ldo)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * J.npos) return;
    const long m = t / J.npos;
    const int p = (int)(t % J.npos);
    auto ld = [&](const double *v, uint32_t idx) -> double {
        if (with_x0) return idx == 0 ? 1. : v[m * stride + idx - 1];
        return v[m * stride + idx];
    };
    double acc = 0.;
    for (int e = J.pos_ptr[p]; e < J.pos_ptr[p + 1]; ++e) {
        const Entry en = J.ent[e];
        double q = ld(va, en.jk & 0xffffu);
        if (RANK == 5) q = q * ld(vb, en.jk >> 16) * ld(vc, en.lm);
        acc += q * en.v;
    }
    out[(m * ldo + (J.pos_i[p] - shift)) * ldo + (J.pos_j[p] - shift)] = acc;
}

}  // namespace qgsb

using namespace qgsb;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *qgsb_last_error(void) { return g_err; }
const char *qgsb_version(void) { return "qgsb 0.1 (sm_100a)"; }
long qgsb_launch_count(void)
{
    long total = 0;
    for (int g = 0; g < MAX_DEVICES; ++g) total += g_ctx[g].launches;
    return total;
}

int qgsb_init(int device)
{
    QGSB_API_BEGIN
    // device < 0 on an initialised library keeps the configuration (start() of every integrator passes -1)
    if (device < 0 && g_slots > 0 && g_ctx[0].ready) {
        QGSB_CUDA(cudaSetDevice(g_ctx[0].device));
        return 0;
    }
    init_device(device);
    QGSB_API_END
}

int qgsb_set_devices(int n_devices, const int *devices)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(n_devices >= 1 && n_devices <= MAX_DEVICES && devices != nullptr, "need 1..%d device ordinals",
                 MAX_DEVICES);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    QGSB_REQUIRE(e == cudaSuccess && count > 0, "no CUDA device available (%s); libqgsb has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> devs(devices, devices + n_devices);
    for (int q = 0; q < n_devices; ++q) {
        QGSB_REQUIRE(devs[q] >= 0 && devs[q] < count, "device %d requested but only %d visible", devs[q], count);
        for (int r = 0; r < q; ++r) QGSB_REQUIRE(devs[r] != devs[q], "device %d listed twice", devs[q]);
    }
    init_devices(devs);
    QGSB_API_END
}

int qgsb_device_count(void)
{
    QGSB_API_LOCK
    return g_slots;
}

void qgsb_shutdown(void)
{
    QGSB_API_LOCK
    close_all();
}

int qgsb_set_stream(void *cuda_stream)
{
    QGSB_API_BEGIN
    ensure_init();
    ctx().stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx().own_stream;
    QGSB_API_END
}

int qgsb_synchronize(void)
{
    QGSB_API_BEGIN
    ensure_init();
    QGSB_CUDA(cudaStreamSynchronize(ctx().stream));
    QGSB_API_END
}

int qgsb_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem)
{
    QGSB_API_BEGIN
    ensure_init();
    Context &c = ctx();
    if (device) *device = c.device;
    if (sm_count) *sm_count = c.sm_count;
    if (cc_major) *cc_major = c.cc_major;
    if (cc_minor) *cc_minor = c.cc_minor;
    if (total_mem) *total_mem = c.total_mem;
    QGSB_API_END
}

int qgsb_load_plugin(const char *path)
{
    QGSB_API_BEGIN
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    QGSB_REQUIRE(h != nullptr, "dlopen(%s) failed: %s", path, dlerror());
    typedef const SpecKernels *(*entry_fn)(void);
    entry_fn entry = (entry_fn)dlsym(h, "qgsb_plugin_kernels");
    QGSB_REQUIRE(entry != nullptr, "%s does not export qgsb_plugin_kernels", path);
    const SpecKernels *k = entry();
    QGSB_REQUIRE(k != nullptr, "%s returned no kernel table", path);
    QGSB_REQUIRE(k->abi == QGSB_SPEC_ABI,
                 "%s was built against kernel-table layout %u, this library uses %u: delete it (it is a cache) and let "
                 "qgs_b200.codegen rebuild it", path, k->abi, (unsigned)QGSB_SPEC_ABI);
    QGSB_REQUIRE((k->rk_chain != nullptr || k->tangent != nullptr), "%s returned an empty kernel table",
                 path);
    register_spec(k);
    QGSB_API_END
}

}  // extern "C"

// builds a handle on the calling thread's device
static qgsb_tensor *tensor_build(int ndim, int rank, long nnz, const int32_t *coo, const double *val, long jnnz,
                                 const int32_t *jcoo, const double *jval)
{
    HostTensor h = prepare_vec(ndim + 1, rank, nnz, coo, val);
    HostJac j = prepare_mat(ndim + 1, rank, jnnz, jcoo, jval, false);
    qgsb_tensor *t = new qgsb_tensor();
    t->device = ctx().device;
    try {
        to_device(t->d_ent, h.ent);
        to_device(t->d_row_ptr, h.row_ptr);
        to_device(t->d_jent, j.ent);
        to_device(t->d_pos_ptr, j.pos_ptr);
        to_device(t->d_pos_i, j.pos_i);
        to_device(t->d_pos_j, j.pos_j);
        to_device(t->d_jrow_ptr, j.row_ptr);
        to_device(t->d_jcol_ptr, j.col_ptr);
        to_device(t->d_jcol_perm, j.col_perm);
    } catch (...) {
        delete t;
        throw;
    }
    t->nnz_in = nnz;
    t->jnnz_in = jnnz;
    t->view.n = ndim;
    t->view.rank = rank;
    t->view.nnz = (int)nnz;
    t->view.ent = t->d_ent.p;
    t->view.row_ptr = t->d_row_ptr.p;
    t->view.jac.npos = (int)j.pos_i.size();
    t->view.jac.pos_ptr = t->d_pos_ptr.p;
    t->view.jac.ent = t->d_jent.p;
    t->view.jac.pos_i = t->d_pos_i.p;
    t->view.jac.pos_j = t->d_pos_j.p;
    t->view.jac.row_ptr = t->d_jrow_ptr.p;
    t->view.jac.col_ptr = t->d_jcol_ptr.p;
    t->view.jac.col_perm = t->d_jcol_perm.p;
    t->hash = tensor_hash(ndim, rank, nnz, h.coo_sorted.data(), h.val_sorted.data());
    t->coo_sorted.swap(h.coo_sorted);
    t->val_sorted.swap(h.val_sorted);
    t->spec = find_spec(t->hash);
    if (t->spec && (t->spec->n != ndim || t->spec->rank != rank || t->spec->nnz != (int)nnz)) t->spec = nullptr;
    t->h_pos_i = j.pos_i;
    t->h_pos_j = j.pos_j;
    t->h_ent = h.ent;
    t->h_row_ptr = h.row_ptr;
    t->h_jent = j.ent;
    t->h_pos_ptr = j.pos_ptr;
    t->jac_matches_spec = jacobian_matches(t);
    return t;
}

namespace qgsb {

// The handle the calling thread's device works with: the caller's handle on its own device, else a replica built on
// first use from the arrays the handle was created from (a few hundred KB; SURVEY.md section 8e: "tensor replicated
// to every GPU").  Replicas follow the handle's specialisation switches and die with it.
const qgsb_tensor *tensor_here(const qgsb_tensor *t)
{
    const int device = ctx().device;
    if (t->device == device) return t;
    std::lock_guard<std::mutex> guard(t->replica_mutex);
    auto it = t->replicas.find(device);
    if (it == t->replicas.end()) {
        qgsb_tensor *r = tensor_build(t->ndim_in, t->view.rank, t->nnz_in, t->in_coo.data(), t->in_val.data(), t->jnnz_in,
                                      t->in_jcoo.data(), t->in_jval.data());
        it = t->replicas.emplace(device, r).first;
    }
    qgsb_tensor *r = it->second;
    r->use_spec = t->use_spec;
    if (r->spec != t->spec) {          // a module was loaded after the replica was made
        r->spec = t->spec;
        r->jac_matches_spec = t->jac_matches_spec;
    }
    return r;
}

}  // namespace qgsb

extern "C" {

int qgsb_tensor_create(int ndim, int rank, long nnz, const int32_t *coo, const double *val, long jnnz,
                       const int32_t *jcoo, const double *jval, qgsb_tensor **out)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(out != nullptr, "null output handle");
    QGSB_REQUIRE(ndim >= 1, "ndim must be positive");
    ensure_init();
    qgsb_tensor *t = tensor_build(ndim, rank, nnz, coo, val, jnnz, jcoo, jval);
    t->ndim_in = ndim;
    t->in_coo.assign(coo, coo + (size_t)nnz * rank);
    t->in_val.assign(val, val + nnz);
    if (jnnz > 0) {
        t->in_jcoo.assign(jcoo, jcoo + (size_t)jnnz * rank);
        t->in_jval.assign(jval, jval + jnnz);
    }
    *out = t;
    QGSB_API_END
}

int qgsb_tensor_hash(int ndim, int rank, long nnz, const int32_t *coo, const double *val, uint64_t *hash)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(hash != nullptr, "null output");
    HostTensor h = prepare_vec(ndim + 1, rank, nnz, coo, val);
    *hash = tensor_hash(ndim, rank, nnz, h.coo_sorted.data(), h.val_sorted.data());
    QGSB_API_END
}

void qgsb_tensor_destroy(qgsb_tensor *t)
{
    if (!t) return;
    QGSB_API_LOCK
    delete t;       // replicas included; every array is freed on its own device
    if (g_slots > 0 && g_ctx[0].ready) cudaSetDevice(g_ctx[0].device);
}

int qgsb_tensor_info(const qgsb_tensor *t, int *ndim, int *rank, long *nnz, long *jnnz, int *kernel_kind,
                     uint64_t *hash)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(t != nullptr, "null tensor handle");
    if (ndim) *ndim = t->view.n;
    if (rank) *rank = t->view.rank;
    if (nnz) *nnz = t->nnz_in;
    if (jnnz) *jnnz = t->jnnz_in;
    if (kernel_kind) *kernel_kind = (t->spec && t->use_spec && t->spec->rk_chain) ? 2 : (t->view.n > QGSB_G1_MAX_NDIM ? 1 : 0);
    if (hash) *hash = t->hash;
    QGSB_API_END
}

int qgsb_tensor_has_tangent(const qgsb_tensor *t)
{
    return t && t->spec && t->use_spec && t->spec->tangent && t->jac_matches_spec ? 1 : 0;
}

int qgsb_tensor_use_specialised(qgsb_tensor *t, int enable)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(t != nullptr, "null tensor handle");
    t->use_spec = enable != 0;
    if (enable) {  // pick up modules registered after the handle was created (qgsb_load_plugin)
        const SpecKernels *k = find_spec(t->hash);
        if (k && k->n == t->view.n && k->rank == t->view.rank && k->nnz == (int)t->nnz_in) t->spec = k;
        t->jac_matches_spec = jacobian_matches(t);
    }
    QGSB_API_END
}

// ---- raw contractions --------------------------------------------------------------------------
// sparse_mul2/3/4/5 take the tensor as host arrays on every call.  A user loop that calls them with the same tensor
// (what the reference's own f / Df closures do, tendencies.py:98-121) must not pay for sorting, four cudaMallocs and
// the upload each time: the prepared device copy is kept in a small cache keyed by the CONTENT of (coo, val) -- a
// 64-bit FNV-1a over the words of both arrays, plus rank, length and device -- and the vectors go through the
// context's scratch pool.  The arrays may be freed or rewritten between calls: only their content is looked at.
}  // extern "C"

namespace {

struct RawTensor {
    uint64_t key = 0;
    int rank = 0, n1 = 0, device = -1, npos = 0;
    long nnz = 0;
    bool mat = false;
    uint64_t stamp = 0;
    DevBuf<Entry> ent;
    DevBuf<int> a, b, c;          // vec: row_ptr; mat: pos_ptr, pos_i, pos_j
};

std::vector<std::unique_ptr<RawTensor>> raw_cache;
uint64_t raw_clock = 0;
constexpr size_t RAW_CACHE_ENTRIES = 16;

uint64_t content_key(int rank, int n1, long nnz, bool mat, const int32_t *coo, const double *val)
{
    uint64_t h = 1469598103934665603ULL;
    auto mix = [&](uint64_t w) { h = (h ^ w) * 1099511628211ULL; };
    mix((uint64_t)rank);
    mix((uint64_t)n1);
    mix((uint64_t)nnz);
    mix(mat ? 1 : 0);
    const size_t words = (size_t)nnz * rank;
    for (size_t q = 0; q + 1 < words; q += 2) mix((uint64_t)(uint32_t)coo[q] | ((uint64_t)(uint32_t)coo[q + 1] << 32));
    if (words & 1) mix((uint64_t)(uint32_t)coo[words - 1]);
    for (long q = 0; q < nnz; ++q) {
        uint64_t w;
        memcpy(&w, val + q, sizeof(w));
        mix(w);
    }
    return h;
}

RawTensor &raw_tensor(int rank, int n1, long nnz, bool mat, const int32_t *coo, const double *val)
{
    int device = 0;
    QGSB_CUDA(cudaGetDevice(&device));
    const uint64_t key = content_key(rank, n1, nnz, mat, coo, val);
    for (auto &e : raw_cache)
        if (e->key == key && e->rank == rank && e->n1 == n1 && e->nnz == nnz && e->mat == mat && e->device == device) {
            e->stamp = ++raw_clock;
            return *e;
        }
    if (raw_cache.size() >= RAW_CACHE_ENTRIES) {
        size_t oldest = 0;
        for (size_t q = 1; q < raw_cache.size(); ++q)
            if (raw_cache[q]->stamp < raw_cache[oldest]->stamp) oldest = q;
        raw_cache.erase(raw_cache.begin() + oldest);
    }
    std::unique_ptr<RawTensor> e(new RawTensor);
    e->key = key;
    e->rank = rank;
    e->n1 = n1;
    e->nnz = nnz;
    e->mat = mat;
    e->device = device;
    e->stamp = ++raw_clock;
    if (mat) {
        HostJac h = prepare_mat(n1, rank, nnz, coo, val, true);
        to_device(e->ent, h.ent);
        to_device(e->a, h.pos_ptr);
        to_device(e->b, h.pos_i);
        to_device(e->c, h.pos_j);
        e->npos = (int)h.pos_i.size();
    } else {
        HostTensor h = prepare_vec(n1, rank, nnz, coo, val);
        to_device(e->ent, h.ent);
        to_device(e->a, h.row_ptr);
    }
    raw_cache.push_back(std::move(e));
    return *raw_cache.back();
}

}  // namespace

// called when the contexts are closed (qgsb_shutdown, a change of the device list)
namespace qgsb {
void raw_cache_clear() { raw_cache.clear(); }
}  // namespace qgsb

extern "C" {

static void raw_mulvec(int rank, long nnz, const int32_t *coo, const double *val, int n1, const double *va,
                       const double *vb, const double *vc, const double *vd, double *res)
{
    ensure_init();
    cudaStream_t s = ctx().stream;
    RawTensor &R = raw_tensor(rank, n1, nnz, false, coo, val);
    PoolBuf<double> d_v((size_t)5 * n1);
    double *d_out = d_v.p + (size_t)4 * n1;
    const double *vs[4] = {va, vb, vc ? vc : va, vd ? vd : va};
    for (int q = 0; q < (rank == 5 ? 4 : 2); ++q)
        QGSB_CUDA(cudaMemcpyAsync(d_v.p + (size_t)q * n1, vs[q], sizeof(double) * n1, cudaMemcpyHostToDevice, s));
    TensorView T;
    T.n = n1 - 1;
    T.rank = rank;
    T.nnz = (int)nnz;
    T.ent = R.ent.p;
    T.row_ptr = R.a.p;
    const int threads = 128, blocks = (n1 + threads - 1) / threads;
    if (rank == 5)
        mulvec_kernel<5><<<blocks, threads, 0, s>>>(T, 1, n1, d_v.p, d_v.p + n1, d_v.p + 2 * n1, d_v.p + 3 * n1, 0, 0,
                                                    d_out, 0, n1);
    else
        mulvec_kernel<3><<<blocks, threads, 0, s>>>(T, 1, n1, d_v.p, d_v.p + n1, d_v.p, d_v.p, 0, 0, d_out, 0, n1);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
    QGSB_CUDA(cudaMemcpyAsync(res, d_out, sizeof(double) * n1, cudaMemcpyDeviceToHost, s));
    QGSB_CUDA(cudaStreamSynchronize(s));
}

static void raw_mulmat(int rank, long nnz, const int32_t *coo, const double *val, int n1, const double *va,
                       const double *vb, const double *vc, double *res)
{
    ensure_init();
    cudaStream_t s = ctx().stream;
    RawTensor &R = raw_tensor(rank, n1, nnz, true, coo, val);
    PoolBuf<double> d_v((size_t)3 * n1 + (size_t)n1 * n1);
    double *d_out = d_v.p + (size_t)3 * n1;
    const double *vs[3] = {va, vb ? vb : va, vc ? vc : va};
    for (int q = 0; q < (rank == 5 ? 3 : 1); ++q)
        QGSB_CUDA(cudaMemcpyAsync(d_v.p + (size_t)q * n1, vs[q], sizeof(double) * n1, cudaMemcpyHostToDevice, s));
    QGSB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n1 * n1, s));
    JacView J;
    J.npos = R.npos;
    J.pos_ptr = R.a.p;
    J.ent = R.ent.p;
    J.pos_i = R.b.p;
    J.pos_j = R.c.p;
    if (J.npos > 0) {
        const int threads = 128, blocks = (J.npos + threads - 1) / threads;
        if (rank == 5)
            mulmat_kernel<5><<<blocks, threads, 0, s>>>(J, 1, d_v.p, d_v.p + n1, d_v.p + 2 * n1, 0, 0, d_out, 0, n1);
        else
            mulmat_kernel<3><<<blocks, threads, 0, s>>>(J, 1, d_v.p, d_v.p, d_v.p, 0, 0, d_out, 0, n1);
        count_launch();
        QGSB_CUDA(cudaGetLastError());
    }
    QGSB_CUDA(cudaMemcpyAsync(res, d_out, sizeof(double) * n1 * n1, cudaMemcpyDeviceToHost, s));
    QGSB_CUDA(cudaStreamSynchronize(s));
}

int qgsb_sparse_mul3(long nnz, const int32_t *coo, const double *val, int n1, const double *vec_a,
                     const double *vec_b, double *res)
{
    QGSB_API_BEGIN
    raw_mulvec(3, nnz, coo, val, n1, vec_a, vec_b, nullptr, nullptr, res);
    QGSB_API_END
}

int qgsb_sparse_mul5(long nnz, const int32_t *coo, const double *val, int n1, const double *vec_a,
                     const double *vec_b, const double *vec_c, const double *vec_d, double *res)
{
    QGSB_API_BEGIN
    raw_mulvec(5, nnz, coo, val, n1, vec_a, vec_b, vec_c, vec_d, res);
    QGSB_API_END
}

int qgsb_sparse_mul2(long nnz, const int32_t *coo, const double *val, int n1, const double *vec, double *res)
{
    QGSB_API_BEGIN
    raw_mulmat(3, nnz, coo, val, n1, vec, nullptr, nullptr, res);
    QGSB_API_END
}

int qgsb_sparse_mul4(long nnz, const int32_t *coo, const double *val, int n1, const double *vec_a,
                     const double *vec_b, const double *vec_c, double *res)
{
    QGSB_API_BEGIN
    raw_mulmat(5, nnz, coo, val, n1, vec_a, vec_b, vec_c, res);
    QGSB_API_END
}

// ---- f / Df closures -----------------------------------------------------------------------------
int qgsb_tendencies(const qgsb_tensor *t, long N, const double *x, double *out)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && x && out, "null argument");
    QGSB_REQUIRE(N >= 0, "negative member count");
    if (N == 0) return 0;
    ensure_init();
    cudaStream_t s = ctx().stream;
    const int n = t->view.n;
    if (ctx().h_stage && (size_t)N * n <= Context::STAGE_IN) {
        // a handful of states (the reference's f(t, x) on one state, tendencies.py:98-112): through the mapped
        // staging area, one launch and one synchronisation, no copies
        double *hx = ctx().h_stage, *ho = hx + Context::STAGE_IN;
        const long total = N * n;
        memcpy(hx, x, sizeof(double) * total);
        const int threads = 128;
        const unsigned blocks = (unsigned)((total + threads - 1) / threads);
        if (t->view.rank == 5)
            mulvec_kernel<5><<<blocks, threads, 0, s>>>(t->view, N, n + 1, hx, hx, hx, hx, n, 1, ho, 1, n);
        else
            mulvec_kernel<3><<<blocks, threads, 0, s>>>(t->view, N, n + 1, hx, hx, hx, hx, n, 1, ho, 1, n);
        count_launch();
        QGSB_CUDA(cudaGetLastError());
        QGSB_CUDA(cudaStreamSynchronize(s));
        memcpy(out, ho, sizeof(double) * total);
        return 0;
    }
    PoolBuf<double> d_x((size_t)N * n), d_out((size_t)N * n);
    d_x.upload(x, (size_t)N * n, s);
    if (t->spec && t->use_spec && t->spec->tendencies && N >= 512) {
        // batches: the generated straight-line kernel (state in registers) on the tiled layout
        const long ld = round_up(N, TILE);
        PoolBuf<double> d_xs((size_t)n * ld), d_os((size_t)n * ld);
        launch_aos_to_soa(d_x.p, d_xs.p, N, n, ld);
        QGSB_CUDA(t->spec->tendencies(d_xs.p, d_os.p, ld, N, s));
        count_launch();
        launch_soa_to_aos(d_os.p, d_out.p, N, n, ld);
        d_out.download(out, (size_t)N * n, s);
        QGSB_CUDA(cudaStreamSynchronize(s));
        return 0;
    }
    const long total = N * n;
    const int threads = 128;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (t->view.rank == 5)
        mulvec_kernel<5><<<blocks, threads, 0, s>>>(t->view, N, n + 1, d_x.p, d_x.p, d_x.p, d_x.p, n, 1, d_out.p, 1, n);
    else
        mulvec_kernel<3><<<blocks, threads, 0, s>>>(t->view, N, n + 1, d_x.p, d_x.p, d_x.p, d_x.p, n, 1, d_out.p, 1, n);
    count_launch();
    QGSB_CUDA(cudaGetLastError());
    d_out.download(out, (size_t)N * n, s);
    QGSB_CUDA(cudaStreamSynchronize(s));
    QGSB_API_END
}

int qgsb_jacobian(const qgsb_tensor *t, long N, const double *x, double *out)
{
    QGSB_API_BEGIN
    QGSB_REQUIRE(t && x && out, "null argument");
    QGSB_REQUIRE(N >= 0, "negative member count");
    if (N == 0) return 0;
    ensure_init();
    cudaStream_t s = ctx().stream;
    const int n = t->view.n;
    const long total = N * t->view.jac.npos;
    if (ctx().h_stage && (size_t)N * n <= Context::STAGE_IN && (size_t)N * n * n <= Context::STAGE_OUT) {
        // one state or a few (Df(t, x), tendencies.py:114-121): through the mapped staging area, see qgsb_tendencies
        double *hx = ctx().h_stage, *ho = hx + Context::STAGE_IN;
        memcpy(hx, x, sizeof(double) * N * n);
        memset(ho, 0, sizeof(double) * N * n * n);
        if (total > 0) {
            const int threads = 128;
            const unsigned blocks = (unsigned)((total + threads - 1) / threads);
            if (t->view.rank == 5)
                mulmat_kernel<5><<<blocks, threads, 0, s>>>(t->view.jac, N, hx, hx, hx, n, 1, ho, 1, n);
            else
                mulmat_kernel<3><<<blocks, threads, 0, s>>>(t->view.jac, N, hx, hx, hx, n, 1, ho, 1, n);
            count_launch();
            QGSB_CUDA(cudaGetLastError());
            QGSB_CUDA(cudaStreamSynchronize(s));
        }
        memcpy(out, ho, sizeof(double) * N * n * n);
        return 0;
    }
    PoolBuf<double> d_x((size_t)N * n), d_out((size_t)N * n * n);
    d_x.upload(x, (size_t)N * n, s);
    QGSB_CUDA(cudaMemsetAsync(d_out.p, 0, sizeof(double) * N * n * n, s));
    if (total > 0) {
        const int threads = 128;
        const unsigned blocks = (unsigned)((total + threads - 1) / threads);
        if (t->view.rank == 5)
            mulmat_kernel<5><<<blocks, threads, 0, s>>>(t->view.jac, N, d_x.p, d_x.p, d_x.p, n, 1, d_out.p, 1, n);
        else
            mulmat_kernel<3><<<blocks, threads, 0, s>>>(t->view.jac, N, d_x.p, d_x.p, d_x.p, n, 1, d_out.p, 1, n);
        count_launch();
        QGSB_CUDA(cudaGetLastError());
    }
    d_out.download(out, (size_t)N * n * n, s);
    QGSB_CUDA(cudaStreamSynchronize(s));
    QGSB_API_END
}

}  // extern "C"
