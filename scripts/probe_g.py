"""A/B of the members per block of the packed tangent kernels (QGSB_PACK_G is read at every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
_lib.init(0)
for g in ("7", "4", "3", "2"):
    os.environ["QGSB_PACK_G"] = g
    print("== QGSB_PACK_G=%s" % g, flush=True)
    tgls("maooam36", 8192, 50)
    lyap("maooam36", 8192, 20, 80)
