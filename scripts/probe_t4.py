"""T4 / dynT RK4 throughput on the generated kernels at a few ensemble sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts import bench_suite as bs
_lib.init(0)
for name, N, steps in (("T4", 131072, 50), ("T4", 37888, 50), ("T4", 151552, 50), ("T4", 262144, 20), ("dynT", 1048576, 200)):
    r = bs.rk(name, N, steps)
    print(name, N, steps, "%.3f ms  %.4g member-steps/s  %.2f TFLOP/s" % (r["ms"], r["member_steps_per_s"], r["tflops_algorithmic"]), flush=True)
