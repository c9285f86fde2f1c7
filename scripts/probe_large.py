"""Large-basis (6x6 atmosphere, 228 variables) RK4 throughput: G3 (thread per member) against G2 (warp per member)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from scripts.perf_probe import run  # noqa: E402

_lib.init(0)
os.environ["QGSB_RK_LARGE"] = "g2"
run("atm6x6", 1 << 12, 10, spec=False)
os.environ["QGSB_RK_LARGE"] = "g3"
run("atm6x6", 1 << 12, 10, spec=False)
run("atm6x6", 148 * 96, 10, spec=False)
run("atm6x6", 1 << 16, 20, spec=False)
