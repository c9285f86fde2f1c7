"""A/B of the member-fastest column layout of the QR phase of the packed Benettin kernel (QGSB_QR_REMAP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap
_lib.init(0)
for v in ("0", "1"):
    os.environ["QGSB_QR_REMAP"] = v
    print("== QGSB_QR_REMAP=%s" % v, flush=True)
    lyap("maooam36", 8192, 20, 80)
    lyap("maooam36", 8192, 20, 80, m=10)
    lyap("rp", 8192, 20, 80)
    lyap("dynT", 2048, 10, 40)
