"""T4 / MAOOAM-36 fused RK kernel with two blocks (default) or one block per SM (QGSB_RK_ONE_BLOCK=1, read once per
process): is the T4 kernel bound by instruction fetch per scheduler (throughput unchanged with half the warps) or per warp
(throughput halves)?   python scripts/probe_t4_blocks.py"""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys; sys.path.insert(0, %r)\n"
        "from qgs_b200 import _lib; _lib.init(0)\n"
        "from scripts import bench_suite as bs\n"
        "for name, N, steps in (('T4', 148 * 2 * 128 * 4, 50), ('maooam36', 1 << 20, 200), ('dynT', 1 << 20, 100)):\n"
        "    r = bs.rk(name, N, steps); print(name, '%%.4g member-steps/s' %% r['member_steps_per_s'], flush=True)\n" % REPO)
for one in ("0", "1"):
    print("== QGSB_RK_ONE_BLOCK=%s" % one, flush=True)
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, QGSB_RK_ONE_BLOCK=one))
