"""ncu target: a small ensemble (default 148 members of the 6x6 model) on the latency-regime kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from qgs_b200.functions.tendencies import tendencies_from_tensor
from qgs_b200.integrators.integrator import RungeKuttaIntegrator
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "atm6x6"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
integ = RungeKuttaIntegrator()
integ.set_func(f)
ic = np.random.default_rng(0).random((N, f.ndim)) * 0.01
for _ in range(2):
    integ.integrate(0., 10., 0.1, ic=ic, write_steps=0)
