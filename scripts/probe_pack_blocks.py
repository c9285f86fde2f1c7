"""A/B of the block shape of the packed tangent / Benettin kernels: libraries built with
QGSB_BUILD_TAG=<tag> QGSB_NVCC_EXTRA="-DQGSB_PACK_THREADS=<t> -DQGSB_PACK_BLOCKS=<b>" python -m qgs_b200.build
are run one after the other (QGSB_LIB selects the library at import time, so each variant is its own process).

    python scripts/probe_pack_blocks.py [tag ...]        ("" = the product library)
"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BODY = """
import sys; sys.path.insert(0, %r)
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
_lib.init(0)
tgls("maooam36", 8192, 50)
lyap("maooam36", 8192, 20, 80)
lyap("maooam36", 8192, 20, 80, m=10)
lyap("rp", 8192, 20, 80)
lyap("dynT", 2048, 10, 40)
import os
os.environ["QGSB_QR_ROLLED"] = "1"
print("rolled QR:")
lyap("maooam36", 8192, 20, 80)
lyap("rp", 8192, 20, 80)
""" % REPO

for tag in (sys.argv[1:] or [""]):
    env = dict(os.environ)
    if tag and tag != "default":
        env["QGSB_LIB"] = os.path.join(REPO, "qgs_b200", "libqgsb_%s.so" % tag)
    print("== %s" % (tag or "default"), flush=True)
    subprocess.run([sys.executable, "-c", BODY], env=env, check=False)
