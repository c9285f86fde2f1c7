"""A/B of the unrolled and the rolled Householder QR inside the packed Benettin kernel (QGSB_QR_ROLLED)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts import bench_suite as bs
_lib.init(0)
for rolled in ("0", "1", "0", "1"):
    os.environ["QGSB_QR_ROLLED"] = rolled
    for name, N, steps, m in (("maooam36", 8192, 100, 36), ("maooam36", 8192, 100, 10), ("rp", 8192, 100, 20), ("dynT", 4096, 50, 38)):
        r = bs.tangent(name, N, steps, m, True)
        print("rolled=%s %-9s m=%2d  %8.3f ms  %.4g member-steps/s" % (rolled, name, m, r["ms"], r["member_steps_per_s"]), flush=True)
