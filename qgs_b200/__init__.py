"""qgs_b200 -- B200 (sm_100a) ensemble integrator for the qgs spectral climate model hot path.

Mirror of the reference's interface for the path BASELINE.json names:

* ``qgs_b200.functions.sparse_mul``      <- qgs/functions/sparse_mul.py
* ``qgs_b200.functions.tendencies``      <- qgs/functions/tendencies.py
* ``qgs_b200.integrators.integrate``     <- qgs/integrators/integrate.py
* ``qgs_b200.integrators.integrator``    <- qgs/integrators/integrator.py
* ``qgs_b200.toolbox.lyapunov``          <- qgs/toolbox/lyapunov.py

All arithmetic runs in hand-written CUDA kernels behind the C ABI of ``include/qgsb.h``
(``libqgsb.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
