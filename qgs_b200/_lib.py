"""ctypes binding of ``libqgsb.so`` -- the stub a qgs maintainer would add to call the CUDA path.

Every prototype mirrors ``include/qgsb.h``.  There is no fallback: if the shared library has not
been built (``python -m qgs_b200.build`` / ``__graft_entry__.build()``) or no CUDA device is
present, calls raise ``RuntimeError`` instead of computing on the CPU.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QGSB_LIB") or os.path.join(HERE, "libqgsb.so")      # QGSB_LIB: A/B builds

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_long_p = ctypes.POINTER(ctypes.c_long)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

_PROTOTYPES = {
    "qgsb_init": (ctypes.c_int, [ctypes.c_int]),
    "qgsb_set_devices": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "qgsb_device_count": (ctypes.c_int, []),
    "qgsb_set_seed": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_long]),
    "qgsb_shutdown": (None, []),
    "qgsb_set_stream": (ctypes.c_int, [ctypes.c_void_p]),
    "qgsb_last_error": (ctypes.c_char_p, []),
    "qgsb_version": (ctypes.c_char_p, []),
    "qgsb_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 4 + [ctypes.POINTER(ctypes.c_size_t)]),
    "qgsb_launch_count": (ctypes.c_long, []),
    "qgsb_load_plugin": (ctypes.c_int, [ctypes.c_char_p]),
    "qgsb_tensor_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_long, c_int32_p, c_double_p,
                                          ctypes.c_long, c_int32_p, c_double_p, c_void_pp]),
    "qgsb_tensor_destroy": (None, [ctypes.c_void_p]),
    "qgsb_tensor_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                        c_long_p, c_long_p, ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_uint64)]),
    "qgsb_tensor_hash": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_long, c_int32_p, c_double_p,
                                        ctypes.POINTER(ctypes.c_uint64)]),
    "qgsb_tensor_has_tangent": (ctypes.c_int, [ctypes.c_void_p]),
    "qgsb_tensor_use_specialised": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "qgsb_sparse_mul3": (ctypes.c_int, [ctypes.c_long, c_int32_p, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                        c_double_p]),
    "qgsb_sparse_mul5": (ctypes.c_int, [ctypes.c_long, c_int32_p, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_double_p]),
    "qgsb_sparse_mul2": (ctypes.c_int, [ctypes.c_long, c_int32_p, c_double_p, ctypes.c_int, c_double_p, c_double_p]),
    "qgsb_sparse_mul4": (ctypes.c_int, [ctypes.c_long, c_int32_p, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                        c_double_p, c_double_p]),
    "qgsb_tendencies": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, c_double_p]),
    "qgsb_jacobian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, c_double_p]),
    "qgsb_rk_integrate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_long, c_double_p,
                                         ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                         ctypes.c_int, ctypes.c_long, c_double_p, c_double_p]),
    "qgsb_rk_tgls_integrate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int, c_double_p,
                                              ctypes.c_long, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                              c_double_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_long, c_double_p, c_double_p, c_double_p]),
    "qgsb_lyap_benettin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int, ctypes.c_int,
                                          c_double_p, c_double_p, ctypes.c_long, ctypes.c_long, c_double_p, c_long_p,
                                          c_double_p, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                          ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_long, c_double_p,
                                          c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "qgsb_clv_ginelli": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                        ctypes.c_long, ctypes.c_long, ctypes.c_long, c_double_p, c_long_p, c_double_p,
                                        ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                        ctypes.c_double, c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                        c_double_p, c_double_p, c_double_p, c_double_p]),
    "qgsb_clv_subspace_intersect": (ctypes.c_int, [ctypes.c_long, ctypes.c_int, ctypes.c_long, c_double_p, c_double_p,
                                                   c_double_p]),
    "qgsb_ensemble_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_void_pp]),
    "qgsb_ensemble_destroy": (None, [ctypes.c_void_p]),
    "qgsb_ensemble_upload": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "qgsb_ensemble_download": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "qgsb_ensemble_integrate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int, c_double_p,
                                               c_double_p, c_double_p, c_double_p]),
    "qgsb_ensemble_integrate_record": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int,
                                                      c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                                      ctypes.c_long, ctypes.c_void_p, c_double_p]),
    "qgsb_ensemble_integrate_trajectories": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int,
                                                            c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                                            ctypes.c_int, ctypes.c_long, c_double_p, c_double_p]),
    "qgsb_ensemble_moments": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p]),
    "qgsb_ensemble_integrate_moments": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, c_double_p, ctypes.c_int,
                                                       c_double_p, c_double_p, c_double_p, ctypes.c_long,
                                                       ctypes.c_long, c_double_p, c_double_p, c_double_p]),
    "qgsb_ensemble_device_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "qgsb_ensemble_ld": (ctypes.c_long, [ctypes.c_void_p]),
    "qgsb_synchronize": (ctypes.c_int, []),
    "qgsb_fp64_peak": (ctypes.c_int, [c_double_p, c_double_p]),
    "qgsb_dmma_peak": (ctypes.c_int, [c_double_p]),
}

_lib = None


def exported_names():
    """Names ``include/qgsb.h`` declares (used by the CPU test that checks the library exports them)."""
    return sorted(_PROTOTYPES)


def load():
    """Load libqgsb.so (no device needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libqgsb.so is missing (%s): build it with `python -m qgs_b200.build`; "
                               "qgs_b200 has no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libqgsb: " + load().qgsb_last_error().decode("utf-8", "replace"))


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def init(device=-1):
    check(load().qgsb_init(int(device)))


def set_devices(devices):
    """Drive exactly these CUDA devices (the first is the primary one); see ``qgsb_set_devices``."""
    devices = [int(d) for d in devices]
    arr = (ctypes.c_int * len(devices))(*devices)
    check(load().qgsb_set_devices(len(devices), arr))


def device_count():
    """Number of devices the library drives (0 before the first ``init``)."""
    return int(load().qgsb_device_count())


def set_seed(seed, member_offset=0):
    """Seed of the start bases the Benettin kernels draw on the device (``qgsb_set_seed``)."""
    check(load().qgsb_set_seed(ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), int(member_offset)))


def device_info():
    dev, sm, ma, mi = (ctypes.c_int() for _ in range(4))
    mem = ctypes.c_size_t()
    check(load().qgsb_device_info(ctypes.byref(dev), ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi),
                                  ctypes.byref(mem)))
    return {"device": dev.value, "sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}


def launch_count():
    return int(load().qgsb_launch_count())


def fp64_peak():
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(load().qgsb_fp64_peak(ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value


def dmma_peak():
    tf = ctypes.c_double()
    check(load().qgsb_dmma_peak(ctypes.byref(tf)))
    return tf.value
