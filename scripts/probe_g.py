"""A/B of the members per block of the packed tangent kernels (QGSB_PACK_G is read at every launch; the ceiling is
pack::MAX_THREADS / n_vec, i.e. the QGSB_PACK_THREADS the library was built with)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
_lib.init(0)
for g in (sys.argv[1:] or ("9", "8", "7", "3")):
    os.environ["QGSB_PACK_G"] = g
    print("== QGSB_PACK_G=%s" % g, flush=True)
    tgls("maooam36", 8192, 50)
    lyap("maooam36", 8192, 20, 80)
    lyap("maooam36", 8192, 20, 80, m=10)
