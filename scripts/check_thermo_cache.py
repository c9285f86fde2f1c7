"""Device check of the tensor cache on the vertical-velocity diagnostic: the second construction reads both tensors
(tendencies and atmospheric thermodynamic part) from QGSB_TENSOR_CACHE and gives bitwise the same omega term."""
import os, subprocess, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import write_plot_stubs  # noqa: E402
work = tempfile.mkdtemp()
stubs = write_plot_stubs(os.path.join(work, "stubs"))
code = """
import warnings; warnings.filterwarnings('ignore')
import sys, glob, os, numpy as np
from qgs.params.params import QgParams
from qgs.diagnostics.wind import MiddleLayerVerticalVelocity
p = QgParams()
p.set_atmospheric_channel_fourier_modes(2, 2)
p.set_oceanic_basin_fourier_modes(2, 4)
p.set_params({'kd': 0.0290, 'kdp': 0.0290, 'n': 1.5, 'r': 1.e-7, 'h': 136.5, 'd': 1.1e-7})
p.atemperature_params.set_insolation(103.3333, 0)
p.gotemperature_params.set_insolation(310., 0)
out = []
for k in range(2):
    if k == 1:
        import qgs_b200.functions.tendencies as tmod
        def no_build(*a, **kw):
            raise AssertionError('tensor construction ran although the cache holds both tensors')
        tmod._build_reference_tensor = no_build
    diag = MiddleLayerVerticalVelocity(p)
    data = np.random.default_rng(5).random((p.ndim, 64)) * 0.01
    diag.set_data(np.arange(64) * 0.1, data)
    out.append(diag._data.copy())
files = sorted(os.path.basename(f).split('_')[0] for f in glob.glob(os.path.join(os.environ['QGSB_TENSOR_CACHE'], '*.npz')))
assert files == ['atmo', 'tendencies'], files
assert np.array_equal(out[0], out[1])
print('thermo cache ok', out[0].shape)
"""
env = dict(os.environ, PYTHONPATH=os.pathsep.join([stubs, os.path.join(REPO, "overlay"), os.path.join(REPO, "baseline", "_ref")]),
           QGSB_TENSOR_CACHE=os.path.join(work, "cache"))
res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
print(res.stdout[-500:], res.stderr[-1500:])
sys.exit(res.returncode)
