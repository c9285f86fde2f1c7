#!/usr/bin/env python
"""Build the tendencies tensors of the five BASELINE.json configurations with the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_tensors.py [rp maooam36 dynT T4 atm6x6 ...]

For every configuration the reference's own ``create_tendencies(params, return_qgtensor=True)``
(``/root/reference/qgs/functions/tendencies.py:20-130``) is called and the two arrays the hot path
consumes -- ``coo = tensor.coords.T`` / ``val = tensor.data`` and the Jacobian pair
(``tendencies.py:92-96``) -- are written to ``tests/golden/tensor_<name>.npz`` (int16 indices,
float64 values).  These files are the *input format* fixtures; the GPU box has
no /root/reference, so the tests, smoke() and bench.py read these instead of rebuilding the tensors.

The reference imports pydata ``sparse`` and ``pebble`` which are absent from the image; the
stand-ins in ``qgs_b200/compat`` are put on ``sys.path`` for tensor construction only.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("QGS_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qgs_b200", "compat"))
sys.path.insert(0, REFERENCE)
warnings.filterwarnings("ignore", category=SyntaxWarning)


def params_rp():
    """qgs_rp.py:77-83 (Reinhold & Pierrehumbert 20-variable atmosphere with orography)."""
    from qgs.params.params import QgParams
    p = QgParams({'phi0_npi': np.deg2rad(50.) / np.pi, 'hd': 0.1})
    p.set_atmospheric_channel_fourier_modes(2, 2)
    p.ground_params.set_orography(0.2, 1)
    p.atemperature_params.set_thetas(0.2, 0)
    return p


def params_tlad():
    """model_test/test_tlad.py:15-21 (same model, parameters of the TL/AD test)."""
    from qgs.params.params import QgParams
    p = QgParams({'phi0_npi': np.deg2rad(50.) / np.pi, 'hd': 0.3})
    p.set_atmospheric_channel_fourier_modes(2, 2)
    p.ground_params.set_orography(0.4, 1)
    p.atemperature_params.set_thetas(0.2, 0)
    return p


def params_maooam36():
    """qgs_maooam.py:78-92 (MAOOAM 36 variables) -- the headline configuration."""
    from qgs.params.params import QgParams
    p = QgParams()
    p.set_atmospheric_channel_fourier_modes(2, 2)
    p.set_oceanic_basin_fourier_modes(2, 4)
    p.set_params({'kd': 0.0290, 'kdp': 0.0290, 'n': 1.5, 'r': 1.e-7, 'h': 136.5, 'd': 1.1e-7})
    p.atemperature_params.set_params({'eps': 0.7, 'T0': 289.3, 'hlambda': 15.06, })
    p.gotemperature_params.set_params({'gamma': 5.6e8, 'T0': 301.46})
    p.atemperature_params.set_insolation(103.3333, 0)
    p.gotemperature_params.set_insolation(310., 0)
    return p


def params_aotensor_ref():
    """model_test/test_aotensor.py:43-50 -- the parameter set pinned by test_aotensor.ref."""
    from qgs.params.params import QgParams
    p = QgParams({'rr': 287.e0, 'sb': 5.6e-8})
    p.set_atmospheric_channel_fourier_modes(2, 2)
    p.set_oceanic_basin_fourier_modes(2, 4)
    p.set_params({'kd': 0.04, 'kdp': 0.04, 'n': 1.5})
    return p


def _params_notebook_T4(**flags):
    """notebooks/maooam_T4.ipynb (model cells)."""
    from qgs.params.params import QgParams
    p = QgParams({'n': 1.5}, **flags)
    p.set_atmospheric_channel_fourier_modes(2, 2, mode="symbolic")
    p.set_oceanic_basin_fourier_modes(2, 4, mode="symbolic")
    p.set_params({'kd': 0.0290, 'kdp': 0.0290, 'r': 1.e-7, 'h': 136.5, 'd': 1.1e-7})
    p.atemperature_params.set_params({'eps': 0.7, 'hlambda': 15.06})
    p.gotemperature_params.set_params({'gamma': 5.6e8})
    p.atemperature_params.set_insolation(103., 0)
    p.atemperature_params.set_insolation(103., 1)
    p.gotemperature_params.set_insolation(310., 0)
    p.gotemperature_params.set_insolation(310., 1)
    return p


def params_T4():
    return _params_notebook_T4(T4=True)


def params_dynT():
    return _params_notebook_T4(dynamic_T=True)


def params_atm6x6():
    """model_test/test_aotensor_6x6.py:43-48 (6x6 atmosphere + 6x6 ocean, 228 variables)."""
    from qgs.params.params import QgParams
    p = QgParams({'rr': 287.e0, 'sb': 5.6e-8})
    p.set_atmospheric_channel_fourier_modes(6, 6)
    p.set_oceanic_basin_fourier_modes(6, 6)
    p.set_params({'kd': 0.04, 'kdp': 0.04, 'n': 1.5})
    return p


CONFIGS = {
    "rp": params_rp,
    "tlad": params_tlad,
    "maooam36": params_maooam36,
    "aotensor_ref": params_aotensor_ref,
    "dynT": params_dynT,
    "T4": params_T4,
    "atm6x6": params_atm6x6,
}


def build(name):
    from qgs.functions.tendencies import create_tendencies
    t0 = time.time()
    params = CONFIGS[name]()
    f, Df, qgtensor = create_tendencies(params, return_qgtensor=True)
    coo = np.ascontiguousarray(qgtensor.tensor.coords.T)
    val = np.ascontiguousarray(qgtensor.tensor.data)
    jcoo = np.ascontiguousarray(qgtensor.jacobian_tensor.coords.T)
    jval = np.ascontiguousarray(qgtensor.jacobian_tensor.data)
    assert coo.max() < 32767
    out = os.path.join(HERE, "tensor_%s.npz" % name)
    np.savez_compressed(out, ndim=np.int64(params.ndim), rank=np.int64(coo.shape[1]),
                        coo=coo.astype(np.int16), val=val, jcoo=jcoo.astype(np.int16), jval=jval)
    print("%-12s ndim=%d rank=%d nnz=%d jnnz=%d  %.1fs -> %s" %
          (name, params.ndim, coo.shape[1], len(val), len(jval), time.time() - t0, out), flush=True)


if __name__ == "__main__" and "--pins" not in sys.argv:
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        build(n)


def pin_from_ref_files():
    """Convert the reference's own golden files (model_test/test_aotensor*.ref: one line per tensor
    entry, '%.5E' values) into compact npz pins: tests/golden/refpin_<name>.npz."""
    import re
    pat = re.compile(r"^\w+((?:\[\d+\])+) = \s*(\S+)$")
    for ref, out in (("test_aotensor.ref", "refpin_aotensor"), ("test_aotensor_jacobian.ref", "refpin_aotensor_jacobian"),
                     ("test_aotensor_6x6.ref", "refpin_aotensor_6x6")):
        idx, vals = [], []
        for line in open(os.path.join(REFERENCE, "model_test", ref)):
            m = pat.match(line.strip())
            if not m:
                continue
            idx.append([int(t) for t in re.findall(r"\[(\d+)\]", m.group(1))])
            vals.append(float(m.group(2)))
        np.savez_compressed(os.path.join(HERE, out + ".npz"), coo=np.array(idx, dtype=np.int16), val=np.array(vals))
        print("%s: %d entries" % (out, len(vals)))


if __name__ == "__main__" and "--pins" in sys.argv:
    pin_from_ref_files()
