"""A/B of the two re-orthonormalisations inside the packed Benettin kernel (QGSB_QR_ROLLED) and of the dealing of the
columns to threads (QGSB_QR_REMAP).  The round-2 measurements of the two forms that were removed -- reflectors handed over
through progress flags, and the step as two launches -- are kept in profiles/r02_benettin_qr_pipelined_ab.log and
profiles/r02_benettin_split_ab.log; two reflectors per barrier: profiles/r02_benettin_qr_panel2_ab.log."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts import bench_suite as bs
_lib.init(0)
cases = (("maooam36", 8192, 100, 36), ("maooam36", 8192, 100, 10), ("rp", 8192, 100, 20), ("dynT", 4096, 50, 38))
peak = _lib.fp64_peak()
for rep in range(2):
    for rolled, remap in (("0", "1"), ("1", "1"), ("0", "0")):
        os.environ["QGSB_QR_ROLLED"] = rolled
        os.environ["QGSB_QR_REMAP"] = remap
        for name, N, steps, m in cases:
            r = bs.tangent(name, N, steps, m, True)
            print("rolled=%s remap=%s %-9s m=%2d  %8.3f ms  %.4g member-steps/s  %.3f of FP64 peak (%.1f)"
                  % (rolled, remap, name, m, r["ms"], r["member_steps_per_s"], r["tflops_algorithmic"] / peak, peak),
                  flush=True)
