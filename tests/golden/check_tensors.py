#!/usr/bin/env python
"""Rebuild the tensor fixtures with the UNMODIFIED reference on top of the coordinate-based ``sparse`` stand-in
(qgs_b200/compat/sparse) and compare with tests/golden/tensor_*.npz; prints the peak resident memory.

    python tests/golden/check_tensors.py [rp tlad maooam36 aotensor_ref atm6x6 dynT T4]

Build container only (needs /root/reference).  Result of round 2: profiles/r02_coo_rebuild.log -- the five analytic
configurations are bit-identical (atm6x6: 27 770 entries), the two symbolic rank-5 ones agree to 1e-15 relative (numpy
takes another summation path for the strided slices of a dense array than for the contiguous copies used here) and carry
one extra entry of magnitude 3e-20 where the dense arithmetic cancelled to an exact zero; T4 peaks at 238 MB resident
(190 MB of that is Python + numpy + sympy) where a dense (39**5) backing array alone is 722 MB.
"""
import os, sys, time, resource, warnings
NAMES = sys.argv[1:]
import numpy as np
warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/qgs_b200/compat"); sys.path.insert(0, "/root/reference")
import importlib.util
spec = importlib.util.spec_from_file_location("mk", "/root/repo/tests/golden/make_tensors.py")
mk = importlib.util.module_from_spec(spec)
mk.__file__ = "/root/repo/tests/golden/make_tensors.py"
src = open("/root/repo/tests/golden/make_tensors.py").read().replace('if __name__ == "__main__"', 'if False')
exec(compile(src, "make_tensors", "exec"), mk.__dict__)
import sparse
print("sparse from", sparse.__file__)
from qgs.functions.tendencies import create_tendencies
for name in (NAMES or ["rp", "tlad", "maooam36", "aotensor_ref"]):
    t0 = time.time()
    params = mk.CONFIGS[name]()
    f, Df, q = create_tendencies(params, return_qgtensor=True)
    coo = q.tensor.coords.T; val = q.tensor.data; jcoo = q.jacobian_tensor.coords.T; jval = q.jacobian_tensor.data
    z = np.load("/root/repo/tests/golden/tensor_%s.npz" % name)
    same = (np.array_equal(coo, z["coo"]) and np.array_equal(val, z["val"]) and np.array_equal(jcoo, z["jcoo"]) and np.array_equal(jval, z["jval"]))
    A = {tuple(c): v for c, v in zip(coo.tolist(), val.tolist())}
    B = {tuple(c): v for c, v in zip(z["coo"].tolist(), z["val"].tolist())}
    scale = np.abs(z["val"]).max()
    close = all(abs(A.get(k, 0.) - B.get(k, 0.)) <= 1e-14 * max(abs(B.get(k, 0.)), 1e-3 * scale) for k in set(A) | set(B))
    print("%-12s nnz=%d jnnz=%d identical=%s close=%s  %.1fs  peak RSS %.0f MB" % (name, len(val), len(jval), same, close, time.time()-t0, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss/1024), flush=True)
