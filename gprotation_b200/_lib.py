"""ctypes binding of libgprot_b200.so (include/gprot_b200.h).

This is the stub a GProtation maintainer would add next to gprot/model.py to route
GPRotModel.lnlike through the B200 engine.  There is no fallback:
if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import CFUNCTYPE, POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("GPROT_B200_LIB", "libgprot_b200.so"))

THETA_ON_DEVICE = 1
OUT_ON_DEVICE = 2
EXACT_KERNEL = 4

# every symbol include/gprot_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "gprot_abi_version": (c_int, []),
    "gprot_create": (c_int, [c_int, POINTER(c_void_p)]),
    "gprot_destroy": (c_int, [c_void_p]),
    "gprot_last_error": (c_char_p, [c_void_p]),
    "gprot_register_lc": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int)]),
    "gprot_unregister_all": (c_int, [c_void_p]),
    "gprot_unregister_lc": (c_int, [c_void_p, c_int]),
    "gprot_generation": (c_int64, [c_void_p]),
    "gprot_lnlike_batch": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "gprot_lnprior_batch": (c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_int, c_void_p]),
    "gprot_lnpost_batch": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                   c_void_p, c_int, c_void_p, c_void_p, c_uint, c_void_p]),
    "gprot_sync": (c_int, [c_void_p]),
    "gprot_set_cluster": (c_int, [c_void_p, c_int]),
    "gprot_last_cluster": (c_int, [c_void_p]),
    "gprot_group_create": (c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    "gprot_group_destroy": (c_int, [c_void_p]),
    "gprot_group_size": (c_int, [c_void_p]),
    "gprot_group_last_error": (c_char_p, [c_void_p]),
    "gprot_group_register_lc": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int)]),
    "gprot_group_unregister_all": (c_int, [c_void_p]),
    "gprot_group_lnlike_batch": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_uint]),
    "gprot_group_unregister_lc": (c_int, [c_void_p, c_int]),
    "gprot_group_lnpost_batch": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                         c_void_p, c_int, c_void_p, c_void_p, c_uint]),
    "gprot_chain_run_fn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "gprot_chain_run": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_int, c_int, c_int64, c_int,
                                c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gprot_chain_last_error": (c_char_p, []),
    "gprot_last_kernel_ms": (c_int, [c_void_p, POINTER(c_float)]),
    "gprot_launch_count": (c_int64, [c_void_p]),
    "gprot_measure_dmma_peak": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "gprot_debug_read_scratch": (c_int, [c_void_p, c_int, c_size_t, c_size_t, c_void_p]),
    "gprot_debug_read_timing": (c_int, [c_void_p, c_void_p]),
}



class ChainMoves(Structure):
    """gprot_chain_moves (include/gprot_b200.h): mixture weights and parameters of the native ensemble sampler."""
    _fields_ = [(n, c_double) for n in ("w_stretch", "w_de", "w_snooker", "w_kde", "stretch_a", "de_sigma", "de_gamma0", "snooker_gamma", "kde_bw")]


# gprot_lnpost_fn: void (*)(void *user, int64_t n, const double *theta, double *lnprior, double *lnlike)
LNPOST_FN = CFUNCTYPE(None, c_void_p, c_int64, POINTER(c_double), POINTER(c_double), POINTER(c_double))

_lib = None


def load():
    """Load the shared library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). gprotation_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
