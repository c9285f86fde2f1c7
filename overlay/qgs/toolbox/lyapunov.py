"""qgs.toolbox.lyapunov served by the CUDA path (qgs_b200.toolbox.lyapunov)."""
from qgs_b200.toolbox.lyapunov import (LyapunovsEstimator, CovariantLyapunovsEstimator,  # noqa: F401
                                       LyapProcess, ClvProcess)
