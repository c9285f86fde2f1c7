"""qgs.functions.tendencies served by the CUDA path (qgs_b200.functions.tendencies)."""
from qgs_b200.functions.tendencies import (create_tendencies, create_atmo_thermo_tendencies,  # noqa: F401
                                           tendencies_from_tensor, tendencies_from_file, save_tendencies, Tendencies, JacobianTendencies, DeviceTensor)
