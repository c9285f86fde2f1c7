"""ncu target: a short packed Benettin (default) or tangent-linear (argument `tgls`) launch, MAOOAM-36."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
_lib.init(0)
if len(sys.argv) > 1 and sys.argv[1] == "tgls":
    tgls("maooam36", 2048, 10)
elif len(sys.argv) > 1 and sys.argv[1] == "exponents":
    # the bench workload's mode: exponents only, every step but the last on the Cholesky QR
    from scripts.probe_cholqr import run
    os.environ["QGSB_QR_CHOL"] = "1"
    run("maooam36", 2072, 4, 36, vectors=False)          # 2072 = 2 whole waves of 148 blocks x 7 members
elif len(sys.argv) > 1 and sys.argv[1].startswith("m="):
    lyap("maooam36", 4096, 4, 16, m=int(sys.argv[1][2:]))
else:
    lyap("maooam36", 2048, 4, 16)
