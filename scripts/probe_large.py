"""Large-basis (6x6 atmosphere, 228 variables) RK4 throughput: G3 two members per thread / one member per thread / G2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib  # noqa: E402
from scripts.perf_probe import run  # noqa: E402

_lib.init(0)
for mode in (sys.argv[1:] or ["pair", "single", "g2"]):
    if mode == "g2":
        os.environ["QGSB_RK_LARGE"] = "g2"
        run("atm6x6", 1 << 12, 10, spec=False)
        os.environ.pop("QGSB_RK_LARGE")
        continue
    os.environ["QGSB_G3_MODE"] = mode
    print("== G3 mode %s" % mode, flush=True)
    run("atm6x6", 148 * 192, 10, spec=False)
    run("atm6x6", 1 << 16, 20, spec=False)
