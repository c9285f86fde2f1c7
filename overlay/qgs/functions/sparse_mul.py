"""qgs.functions.sparse_mul served by the CUDA path (qgs_b200.functions.sparse_mul)."""
from qgs_b200.functions.sparse_mul import sparse_mul2, sparse_mul3, sparse_mul4, sparse_mul5  # noqa: F401
