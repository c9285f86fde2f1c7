"""qgs.integrators.integrate served by the CUDA path (qgs_b200.integrators.integrate)."""
from qgs_b200.integrators.integrate import (integrate_runge_kutta, integrate_runge_kutta_tgls,  # noqa: F401
                                            _integrate_runge_kutta_jit, _integrate_runge_kutta_tgls_jit,
                                            _zeros_func)
