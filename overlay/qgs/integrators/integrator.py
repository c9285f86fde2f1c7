"""qgs.integrators.integrator served by the CUDA path (qgs_b200.integrators.integrator)."""
from qgs_b200.integrators.integrator import (RungeKuttaIntegrator, RungeKuttaTglsIntegrator,  # noqa: F401
                                             TrajectoryProcess, TglsTrajectoryProcess)
