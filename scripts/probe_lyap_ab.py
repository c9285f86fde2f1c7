"""A/B of the tangent-linear / Benettin kernel variants (QGSB_TGLS_KERNEL is read at every launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from scripts.perf_probe2 import lyap, tgls  # noqa: E402

_lib.init(0)
for variant in (sys.argv[1:] or ("generic", "pack_dense", "pack")):
    os.environ["QGSB_TGLS_KERNEL"] = variant
    print("== variant %s" % variant, flush=True)
    tgls("maooam36", 8192, 50)
    lyap("maooam36", 8192, 20, 80)
    lyap("maooam36", 8192, 20, 80, m=10)
    lyap("maooam36", 2048, 10, 40, mdt=0.02)
    lyap("rp", 8192, 20, 80)
    lyap("dynT", 2048, 10, 40)
