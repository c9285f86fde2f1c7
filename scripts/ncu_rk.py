"""ncu target: a short RK4 launch of one tensor on the throughput kernels.   python scripts/ncu_rk.py T4 37888 3"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from qgs_b200.functions.tendencies import tendencies_from_tensor
from qgs_b200.integrators.integrator import RungeKuttaIntegrator
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.init(0)
name, N, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
z = np.load(os.path.join(REPO, "tests", "golden", "tensor_%s.npz" % name))
f, Df = tendencies_from_tensor(int(z["ndim"]), z["coo"], z["val"], z["jcoo"], z["jval"])
integ = RungeKuttaIntegrator()
integ.set_func(f)
ic = np.random.default_rng(0).random((N, f.ndim)) * 0.01
integ.integrate(0., steps * 0.1, 0.1, ic=ic, write_steps=0)
print("kernel_kind", f.tensor.kernel_kind)
