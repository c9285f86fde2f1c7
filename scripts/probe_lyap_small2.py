import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap
_lib.init(0)
for env in ({}, {"QGSB_QR_REMAP": "0"}, {"QGSB_TGLS_KERNEL": "pack_dense"}):
    os.environ.update(env)
    print("== env", env, flush=True)
    for N in (7, 64, 512, 2048):
        lyap("maooam36", N, 20, 80)
    lyap("maooam36", 64, 20, 80, m=10)
    for k in env: os.environ.pop(k)
