"""A/B of the Benettin loop on the packed kernels: one fused launch (QGSB_BENETTIN_SPLIT=0) against two launches per
step (=1: tangent propagation + stand-alone batched QR), with the three factorisations (QGSB_QR_MODE 0 unrolled with a
block barrier per reflector, 1 rolled, 2 pipelined with flags) and one or two QR blocks per SM (QGSB_QR_BLOCKS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts import bench_suite as bs
_lib.init(0)
cases = (("maooam36", 8192, 100, 36), ("maooam36", 8192, 100, 10), ("rp", 8192, 100, 20), ("dynT", 4096, 50, 38))
configs = [("0", "0", "2"), ("0", "1", "2"), ("1", "0", "2"), ("1", "1", "2"), ("1", "2", "2"), ("1", "0", "1"),
           ("1", "1", "1"), ("1", "2", "1")]
peak = _lib.fp64_peak()
for rep in range(2):
    for split, mode, blocks in configs:
        os.environ["QGSB_BENETTIN_SPLIT"] = split
        os.environ["QGSB_QR_MODE"] = mode
        os.environ["QGSB_QR_BLOCKS"] = blocks
        for name, N, steps, m in cases:
            r = bs.tangent(name, N, steps, m, True)
            print("split=%s qr_mode=%s qr_blocks=%s %-9s m=%2d  %8.3f ms  %.4g member-steps/s  %.3f of FP64 peak (%.1f)"
                  % (split, mode, blocks, name, m, r["ms"], r["member_steps_per_s"], r["tflops_algorithmic"] / peak,
                     peak), flush=True)
