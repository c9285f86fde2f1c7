"""qgs.integrators.statistics served by the CUDA path (qgs_b200.integrators.statistics)."""
from qgs_b200.integrators.statistics import TrajectoriesStatistics  # noqa: F401
