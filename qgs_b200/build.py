"""Build libqgsb.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m qgs_b200.build [--force]

Sources: qgs_b200/csrc/*.cu plus the tensor-specialised kernels that qgs_b200/codegen.py
generates for the canonical configurations (tests/golden/tensor_*.npz) under
qgs_b200/csrc/generated/.  nvcc cross-compiles without a GPU; the .so travels to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "generated")
# QGSB_BUILD_TAG=<tag> builds a variant next to the product library (libqgsb_<tag>.so, objects under _obj_<tag>/) for
# A/B measurements; QGSB_LIB=<path> makes qgs_b200._lib load it.
TAG = os.environ.get("QGSB_BUILD_TAG", "")
LIB = os.path.join(HERE, "libqgsb%s.so" % ("_" + TAG if TAG else ""))
OBJ = os.path.join(HERE, "_obj" + ("_" + TAG if TAG else ""))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + \
    os.environ.get("QGSB_NVCC_EXTRA", "").split()          # e.g. -DQGSB_PACK_THREADS=324 for A/B builds


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libqgsb.so cannot be built")
    return exe


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def compile_object(src, obj, extra=()):
    cmd = [nvcc()] + ARCH + NVCC_FLAGS + list(extra) + ["-I", CSRC, "-c", src, "-o", obj]
    subprocess.check_call(cmd)


def build(force=False, verbose=False, generate=True):
    os.makedirs(OBJ, exist_ok=True)
    if generate:
        from . import codegen
        codegen.generate_canonical(GEN)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + sorted(glob.glob(os.path.join(GEN, "*.cu")))
    only = os.environ.get("QGSB_BUILD_SPECS")              # A/B builds: only these generated modules, e.g. "maooam36,rp"
    if only and TAG:
        keep = {"spec_%s.cu" % n for n in only.split(",")}
        sources = [s for s in sources if os.path.dirname(s) != GEN or os.path.basename(s) in keep]
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    objs = []
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [nvcc()] + ARCH + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-I", CSRC, "-c", src, "-o", obj]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (os.path.basename(src), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if jobs or force or _newer(LIB, objs):
        cmd = [nvcc(), "-shared"] + ARCH + ["-o", LIB] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
