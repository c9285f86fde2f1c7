"""A/B of the re-orthonormalisation inside the packed Benettin kernel: QGSB_QR_MODE 0 (unrolled, one block barrier per
reflector), 1 (rolled), 2 (pipelined: flags instead of barriers), with and without QGSB_QR_REMAP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts import bench_suite as bs
_lib.init(0)
cases = (("maooam36", 8192, 100, 36), ("maooam36", 8192, 100, 10), ("rp", 8192, 100, 20), ("dynT", 4096, 50, 38))
for rep in range(2):
    for mode, remap in (("0", "1"), ("1", "1"), ("2", "1"), ("2", "0")):
        os.environ["QGSB_QR_MODE"] = mode
        os.environ["QGSB_QR_REMAP"] = remap
        for name, N, steps, m in cases:
            r = bs.tangent(name, N, steps, m, True)
            print("qr_mode=%s remap=%s %-9s m=%2d  %8.3f ms  %.4g member-steps/s  %.3f of peak-equivalent"
                  % (mode, remap, name, m, r["ms"], r["member_steps_per_s"],
                     r["tflops_algorithmic"] / 36.4), flush=True)
