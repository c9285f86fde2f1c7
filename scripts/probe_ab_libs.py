"""A/B of two builds of the library on the packed tangent / Benettin kernels (QGSB_LIB is read at import time, so each
build runs in its own process):  python scripts/probe_ab_libs.py <tag> [<tag> ...]   ("default" = the product library)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BODY = """
import sys; sys.path.insert(0, %r)
from qgs_b200 import _lib
from scripts.perf_probe2 import lyap, tgls
from scripts.probe_cholqr import run
_lib.init(0)
tgls("maooam36", 8192, 50)
tgls("rp", 8192, 50)
run("maooam36", 8192, 20, 80, vectors=False)
run("maooam36", 8192, 20, 80, m=10, vectors=False)
run("rp", 8192, 20, 80, vectors=False)
""" % REPO

for tag in (sys.argv[1:] or ["default"]):
    env = dict(os.environ)
    if tag != "default":
        env["QGSB_LIB"] = os.path.join(REPO, "qgs_b200", "libqgsb_%s.so" % tag)
    print("== %s" % tag, flush=True)
    subprocess.run([sys.executable, "-c", BODY], env=env, check=False)
