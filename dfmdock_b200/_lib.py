"""ctypes binding of libdfmdock_b200.so (the C ABI declared in include/dfmdock_b200.h).

There is no CPU fallback: if the CUDA library has not been built, or no sm_100 GPU is present, every call
fails loudly (RuntimeError) -- nothing in this package routes around the CUDA kernels.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdfmdock_b200.so")

WANT_ENERGY = 1 << 0
PRECISION_FP32 = 1 << 1
CLASH_FORCE = 1 << 2
NOISE_ANNEAL = 1 << 3
CENTRE_ALL_ATOMS = 1 << 4
ODE = 1 << 5
GRAPH_GENERIC = 1 << 6
LAST_FUSED = 1 << 7
EDGE_SLOTS = 64

# name -> (restype, argtypes); mirrors include/dfmdock_b200.h one to one
PROTOTYPES = {
    "dfm_create": (c_int, [POINTER(c_void_p), c_int]),
    "dfm_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "dfm_finalize_weights": (c_int, [c_void_p, c_float, c_void_p]),
    "dfm_set_complex": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "dfm_set_schedule": (c_int, [c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "dfm_interface_logits": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dfm_set_receptor_pose": (c_int, [c_void_p, c_void_p, c_void_p]),
    "dfm_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "dfm_edges_per_node": (c_int, [c_void_p]),
    "dfm_score_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64,
                                  c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "dfm_reverse_step": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                 c_float, c_float, c_float, c_void_p, c_uint64, c_uint64, c_uint32, c_uint32, c_void_p]),
    "dfm_randomize_pose": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_uint32,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "dfm_sample": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_float, c_uint32, c_uint64, c_uint64,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dfm_launch_count": (c_uint64, [c_void_p]),
    "dfm_profile_enable": (c_int, [c_void_p, c_int]),
    "dfm_profile_read": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(c_int)]),
    "dfm_debug_read": (c_int64, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "dfm_metrics_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dfm_compute_metrics": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "dfm_transform_atoms": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dfm_last_error": (c_char_p, []),
    "dfm_version": (c_char_p, []),
    "dfm_destroy": (None, [c_void_p]),
}

_lib = None


def load():
    """Load the shared library (symbols only -- needs no GPU).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dfmdock_b200: %s is missing; build it with `python -m dfmdock_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().dfm_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())
