"""Generator of tensor-specialised sm_100a kernels.

The reference evaluates the tendencies by looping over a COO list with run-time indices
(``qgs/functions/sparse_mul.py:76-81``).  On the GPU a run-time index forces the model state into
shared memory, and the contraction is then bound by shared-memory bandwidth at roughly a fifth of
the FP64 issue rate.  For one tensor (ndim, rank, index list, values) this module emits the
contraction as straight-line CUDA C++ in which every index is a literal and every tensor value an
immediate: the state, the next-stage state and the weighted stage sum live in registers, no tensor
data is fetched at run time, and a whole Runge-Kutta integration (all stages, all steps, recording)
is one kernel launch.  Terms of a row that share a coefficient magnitude are factored,
``c * (x_a x_b - x_c x_d + ...)``, which the MAOOAM tensors allow for a third of their entries.

(Keeping the values in ``__constant__`` memory instead would let one module serve every parameter
set of a structure, but ptxas hoists those loop-invariant constant loads out of the stage loop and
spills them to local memory -- measured 255 registers + 1.2 KB of spills for the 20-variable model --
so values are baked in and a module is keyed by structure *and* values.)

``generate_canonical`` writes the modules for the configurations of BASELINE.json (from
``tests/golden/tensor_*.npz``) into ``qgs_b200/csrc/generated``; ``qgs_b200.build`` compiles them
into libqgsb.so.  ``build_plugin`` compiles a module for any other tensor with nvcc at run time
(like numba JIT-compiles the reference's ``f`` on first use) and ``qgsb_load_plugin`` loads it.
"""
import collections
import glob
import os
import shutil
import subprocess

import numpy as np

FNV_OFFSET = 1469598103934665603
FNV_PRIME = 1099511628211
MASK64 = (1 << 64) - 1

# tensors outside these bounds stay on the generic kernels (register budget / code size)
MAX_SPEC_TERMS = 1200
MAX_SPEC_NDIM = 38

HERE = os.path.dirname(os.path.abspath(__file__))


def sort_by_row(coo, val):
    """Stable sort of the entries by first index -- the order libqgsb uses on the device."""
    coo = np.ascontiguousarray(coo, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    order = np.argsort(coo[:, 0], kind="stable")
    return np.ascontiguousarray(coo[order]), np.ascontiguousarray(val[order])


def tensor_hash(ndim, rank, coo_sorted, val_sorted):
    """FNV-1a over (ndim, rank, nnz) int32, the row-sorted coo int32 and the values float64;
    mirrors ``qgsb::tensor_hash`` in csrc/runtime.cu."""
    coo_sorted = np.ascontiguousarray(coo_sorted, dtype=np.int32)
    val_sorted = np.ascontiguousarray(val_sorted, dtype=np.float64)
    head = np.array([ndim, rank, coo_sorted.shape[0]], dtype=np.int32)
    data = np.frombuffer(head.tobytes() + coo_sorted.tobytes() + val_sorted.tobytes(), dtype=np.uint8)
    h = FNV_OFFSET
    for byte in data.tolist():
        h = ((h ^ byte) * FNV_PRIME) & MASK64
    return h


def _lit(v):
    return "(%s)" % float(v).hex()


def _prod(factors):
    return " * ".join("x[%d]" % j for j in factors)


def _row_code(i, terms, out):
    """Straight-line code for ``k`` = row ``i``.  terms: list of (factors tuple, value)."""
    stmts = []
    started = [False]

    def add(a, b):  # k (+)= a * b
        if started[0]:
            stmts.append("k = fma(%s, %s, k);" % (a, b))
        else:
            stmts.append("k = %s * %s;" % (a, b))
            started[0] = True

    consts = [v for f, v in terms if len(f) == 0]
    if consts:
        stmts.append("k = %s;" % _lit(sum(consts)))
        started[0] = True
    for f, v in terms:
        if len(f) == 1:
            add(_lit(v), "x[%d]" % f[0])
    high = [(f, v) for f, v in terms if len(f) >= 2]
    # 1. factor out shared coefficient magnitudes: c * (+- prod +- prod ...)
    by_mag = collections.OrderedDict()
    for f, v in high:
        by_mag.setdefault(abs(v), []).append((f, v))
    singles = []
    for mag, grp in by_mag.items():
        if len(grp) < 2:
            singles += grp
            continue
        first = True
        for f, v in grp:
            sgn = "-" if v < 0 else ""
            if first:
                stmts.append("t = %sx[%d] * %s;" % (sgn, f[0], _prod(f[1:])))
                first = False
            elif len(f) == 2:
                stmts.append("t = fma(%sx[%d], x[%d], t);" % (sgn, f[0], f[1]))
            else:
                stmts.append("t = fma(%sx[%d] * %s, x[%d], t);" % (sgn, f[0], _prod(f[1:-1]), f[-1]))
        add(_lit(mag), "t")
    # 2. remaining quadratic terms: x_j * (sum_k c_k x_k) when a factor is shared
    quads = [(f, v) for f, v in singles if len(f) == 2]
    others = [(f, v) for f, v in singles if len(f) > 2]
    while quads:
        cnt = collections.Counter()
        for f, v in quads:
            for j in set(f):
                cnt[j] += 1
        var, _ = max(cnt.items(), key=lambda kv: (kv[1], -kv[0]))
        grp = [q for q in quads if var in q[0]]
        quads = [q for q in quads if var not in q[0]]
        first = True
        for f, v in grp:
            other = f[1] if f[0] == var else f[0]
            if first:
                stmts.append("t = %s * x[%d];" % (_lit(v), other))
                first = False
            else:
                stmts.append("t = fma(%s, x[%d], t);" % (_lit(v), other))
        add("t", "x[%d]" % var)
    for f, v in others:
        stmts.append("t = %s * %s;" % (_lit(v), _prod(f[:-1])))
        add("t", "x[%d]" % f[-1])
    if not started[0]:
        stmts.append("k = 0.;")
    out.append("{ double k, t; (void)t; " + " ".join(stmts) + " ROW_DONE(%d, k); }" % i)


def emit_source(name, ndim, rank, coo_sorted, val_sorted, plugin=False):
    """CUDA C++ source of the specialised module for one tensor -> (source, hash, fp64_instr_per_f)."""
    coo_sorted = np.ascontiguousarray(coo_sorted, dtype=np.int32)
    val_sorted = np.ascontiguousarray(val_sorted, dtype=np.float64)
    nnz = coo_sorted.shape[0]
    h = tensor_hash(ndim, rank, coo_sorted, val_sorted)
    rows = collections.defaultdict(list)
    for c, v in zip(coo_sorted, val_sorted):
        if c[0] == 0 or v == 0.:
            continue  # row 0 is overwritten by 1 (sparse_mul.py:80)
        rows[int(c[0])].append((tuple(int(j) for j in c[1:] if j != 0), float(v)))
    body = []
    for i in range(1, ndim + 1):
        _row_code(i, rows.get(i, []), body)
    text = "\n".join(body)
    n_fp = text.count("fma(") + text.count(" * ")
    src = _TEMPLATE.replace("@F_BODY@", " \\\n    ".join(body)).replace("@NAME@", name) \
        .replace("@HASH@", "0x%016xULL" % h).replace("@NDIM@", str(ndim)).replace("@RANK@", str(rank)) \
        .replace("@NNZ@", str(nnz)).replace("@NFP@", str(n_fp)) \
        .replace("@REGISTER@", _PLUGIN_TAIL if plugin else _STATIC_TAIL)
    return src, h, n_fp


_STATIC_TAIL = "const qgsb::SpecRegistrar registrar(&kernels);\n\n}  // namespace"
_PLUGIN_TAIL = ("}  // namespace\n\nextern \"C\" __attribute__((visibility(\"default\"))) "
                "const qgsb::SpecKernels *qgsb_plugin_kernels(void) { return &kernels; }")

_TEMPLATE = r'''// GENERATED by qgs_b200/codegen.py -- do not edit.  Tensor "@NAME@":
// ndim @NDIM@, rank @RANK@, @NNZ@ entries, hash @HASH@; @NFP@ FP64 instructions per evaluation of f.
//
// Straight-line replacement of the COO loops of qgs/functions/sparse_mul.py:76-81 / :153-158 inside
// the Runge-Kutta loop of qgs/integrators/integrate.py:205-221.  Indices are literals, so x[], xn[]
// and acc[] are registers; tensor values are immediates.
#include "spec_registry.h"

namespace {

constexpr int NDIM = @NDIM@;
constexpr int TILE = 128;        // members per block, one thread each == tile width of the HBM layout
constexpr int MAX_STAGES = 8;

struct Coef {
    double alpha[MAX_STAGES + 1];  // alpha[i] = a[i][i-1]; alpha[s] = 0
    double beta[MAX_STAGES];
};

// x[0] is the constant 1 of the reference's augmented state (tendencies.py:112); x[1..NDIM] the state.
#define F_BODY \
    @F_BODY@

__global__ void __launch_bounds__(TILE, 2)
rk_chain_kernel(double *__restrict__ y_g, long n_steps, const double *__restrict__ dt_g, int s,
                const __grid_constant__ Coef coef, long write_steps, long n_records, size_t rec_stride,
                double *__restrict__ rec)
{
    __shared__ double ys[NDIM][TILE];
    const int tid = threadIdx.x;
    // this block's tile: NDIM rows of TILE members, contiguous in HBM
    double *yt = y_g + (size_t)blockIdx.x * NDIM * TILE + tid;
    double *rt = rec != nullptr ? rec + (size_t)blockIdx.x * NDIM * TILE + tid : nullptr;
    // volatile: keeps the y reads inside the stage loop (hoisted they would cost NDIM more live doubles)
    const volatile double *ysv = &ys[0][tid];
    double x[NDIM + 1], xn[NDIM + 1], acc[NDIM + 1];
    x[0] = 1.;
    xn[0] = 1.;
#pragma unroll
    for (int i = 1; i <= NDIM; ++i) {
        x[i] = yt[(i - 1) * TILE];
        ys[i - 1][tid] = x[i];
    }
    for (long ti = 0; ti < n_steps; ++ti) {
        const double dt = dt_g[ti];
        if (rt != nullptr && write_steps > 0 && ti % write_steps == 0) {        // integrate.py:210-212
#pragma unroll
            for (int i = 1; i <= NDIM; ++i) rt[(i - 1) * TILE] = x[i];
            rt += rec_stride;
        }
#pragma unroll
        for (int i = 1; i <= NDIM; ++i) acc[i] = 0.;
#pragma unroll 1
        for (int st = 0; st < s; ++st) {
            const double wb = dt * coef.beta[st];        // (dt * b)[st]        integrate.py:218
            const double wa = dt * coef.alpha[st + 1];   // (dt * a[st+1])[st]  integrate.py:216
#define ROW_DONE(i, k)                       \
    acc[i] = fma(wb, (k), acc[i]);           \
    xn[i] = fma(wa, (k), ysv[((i) - 1) * TILE]);
            F_BODY
#undef ROW_DONE
#pragma unroll
            for (int i = 1; i <= NDIM; ++i) x[i] = xn[i];
        }
#pragma unroll
        for (int i = 1; i <= NDIM; ++i) {
            x[i] = ys[i - 1][tid] + acc[i];              // y + (dt b) @ k      integrate.py:218-219
            ys[i - 1][tid] = x[i];
        }
    }
    if (rec != nullptr) {                                                        // integrate.py:221
        double *rl = rec + (size_t)(n_records - 1) * rec_stride + (size_t)blockIdx.x * NDIM * TILE + tid;
#pragma unroll
        for (int i = 1; i <= NDIM; ++i) rl[(i - 1) * TILE] = x[i];
    }
#pragma unroll
    for (int i = 1; i <= NDIM; ++i) yt[(i - 1) * TILE] = x[i];
}

__global__ void __launch_bounds__(TILE)
tendencies_kernel(const double *__restrict__ x_g, double *__restrict__ out_g)
{
    const double *xt = x_g + (size_t)blockIdx.x * NDIM * TILE + threadIdx.x;
    double *ot = out_g + (size_t)blockIdx.x * NDIM * TILE + threadIdx.x;
    double x[NDIM + 1];
    x[0] = 1.;
#pragma unroll
    for (int i = 1; i <= NDIM; ++i) x[i] = xt[(i - 1) * TILE];
#define ROW_DONE(i, k) ot[((i) - 1) * TILE] = (k);
    F_BODY
#undef ROW_DONE
}

cudaError_t rk_chain(double *d_y, long ld, long n_members, long n_steps, const double *d_dt, int s,
                     const double *alpha, const double *beta, long write_steps, long n_records,
                     double *d_rec, int sm_count, cudaStream_t stream)
{
    (void)n_members;
    (void)sm_count;
    if (s < 1 || s > MAX_STAGES || ld % TILE != 0) return cudaErrorInvalidValue;
    Coef coef;
    for (int i = 0; i <= MAX_STAGES; ++i) coef.alpha[i] = (i >= 1 && i < s) ? alpha[i] : 0.;
    for (int i = 0; i < MAX_STAGES; ++i) coef.beta[i] = i < s ? beta[i] : 0.;
    rk_chain_kernel<<<(unsigned)(ld / TILE), TILE, 0, stream>>>(d_y, n_steps, d_dt, s, coef, write_steps, n_records,
                                                               (size_t)NDIM * ld, d_rec);
    return cudaGetLastError();
}

cudaError_t tendencies(const double *d_x, double *d_out, long ld, long n_members, cudaStream_t stream)
{
    (void)n_members;
    if (ld % TILE != 0) return cudaErrorInvalidValue;
    tendencies_kernel<<<(unsigned)(ld / TILE), TILE, 0, stream>>>(d_x, d_out);
    return cudaGetLastError();
}

const qgsb::SpecKernels kernels = {@HASH@, @NDIM@, @RANK@, @NNZ@, "@NAME@", rk_chain, tendencies};
@REGISTER@
'''


def eligible(ndim, rank, nnz):
    return ndim <= MAX_SPEC_NDIM and nnz <= MAX_SPEC_TERMS


def generate_canonical(out_dir, golden_dir=None):
    """Write one module per distinct tensor among tests/golden/tensor_*.npz -> {hash: name}."""
    golden_dir = golden_dir or os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    seen = {}
    for path in sorted(glob.glob(os.path.join(golden_dir, "tensor_*.npz"))):
        name = os.path.basename(path)[len("tensor_"):-len(".npz")]
        z = np.load(path)
        ndim, rank = int(z["ndim"]), int(z["rank"])
        if not eligible(ndim, rank, z["coo"].shape[0]):
            continue
        coo, val = sort_by_row(z["coo"], z["val"])
        src, h, _ = emit_source(name, ndim, rank, coo, val)
        if h in seen:
            continue
        seen[h] = name
        target = os.path.join(out_dir, "spec_%s.cu" % name)
        if not os.path.exists(target) or open(target).read() != src:
            with open(target, "w") as fh:
                fh.write(src)
    keep = {"spec_%s.cu" % n for n in seen.values()}
    for path in glob.glob(os.path.join(out_dir, "spec_*.cu")):
        if os.path.basename(path) not in keep:
            os.remove(path)
    return seen


def plugin_dir():
    d = os.environ.get("QGSB_JIT_DIR") or os.path.join(HERE, "_jit")
    os.makedirs(d, exist_ok=True)
    return d


def build_plugin(ndim, rank, coo, val):
    """Compile (or find in the cache) a specialised module for this tensor.  Returns the path of the
    shared object, or None when the tensor is not eligible or nvcc is unavailable."""
    coo, val = sort_by_row(coo, val)
    if not eligible(ndim, rank, coo.shape[0]):
        return None
    h = tensor_hash(ndim, rank, coo, val)
    target = os.path.join(plugin_dir(), "spec_%016x.so" % h)
    if os.path.exists(target):
        return target
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return None
    src, _, _ = emit_source("jit_%016x" % h, ndim, rank, coo, val, plugin=True)
    cu = target[:-3] + ".cu"
    with open(cu, "w") as fh:
        fh.write(src)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(HERE, "csrc"),
           "-o", target + ".tmp", cu]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (subprocess.CalledProcessError, OSError):
        return None
    os.replace(target + ".tmp", target)
    return target
