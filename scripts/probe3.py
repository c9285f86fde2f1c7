import sys; sys.path.insert(0,'/root/repo')
from qgs_b200 import _lib
from scripts.perf_probe import run
_lib.init(0)
run("T4", 1<<17, 50)
run("T4", 1<<19, 100)
run("dynT", 1<<20, 200)
run("maooam36", 1<<20, 200)
