"""ncu target: a short packed Benettin (default) or tangent-linear (argument `tgls`) launch, MAOOAM-36."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
_lib.init(0)
if len(sys.argv) > 1 and sys.argv[1] == "tgls":
    tgls("maooam36", 2048, 10)
elif len(sys.argv) > 1 and sys.argv[1].startswith("m="):
    lyap("maooam36", 4096, 4, 16, m=int(sys.argv[1][2:]))
else:
    lyap("maooam36", 2048, 4, 16)
