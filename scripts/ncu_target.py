"""Small target for ncu: a few launches of the fused MAOOAM-36 RK4 kernel (same shape as bench.py, fewer steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from scripts.perf_probe import run  # noqa: E402

if __name__ == "__main__":
    _lib.init(0)
    name = sys.argv[1] if len(sys.argv) > 1 else "maooam36"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    spec = (sys.argv[4] != "generic") if len(sys.argv) > 4 else True
    run(name, n, steps, spec=spec, reps=3)
