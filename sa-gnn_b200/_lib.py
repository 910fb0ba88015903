"""ctypes binding of ``libsagnn_b200.so`` (the C ABI declared in include/sagnn_b200.h)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SAGNN_B200_LIB selects another build of the same C ABI (tuning experiments); default = in-tree build
_LIB = os.environ.get("SAGNN_B200_LIB") or os.path.join(_HERE, "lib", "libsagnn_b200.so")
_lib = None

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_szp = ctypes.POINTER(ctypes.c_size_t)
vp = ctypes.c_void_p

STATUS = {0: "OK", 1: "INVALID_ARG", 2: "UNSORTED_INPUT", 3: "CUDA_ERROR", 4: "WORKSPACE_TOO_SMALL",
          5: "NOT_FINALIZED", 6: "OUT_OF_RANGE"}

# name -> (restype, argtypes); must list every symbol of include/sagnn_b200.h
SIGNATURES = {
    "sagnn_last_error": (ctypes.c_char_p, []),
    "sagnn_version": (ctypes.c_char_p, []),
    "sagnn_bucket_events": (ctypes.c_int, [vp, vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int64, ctypes.c_int64, vp, vp, vp, c_i64p, vp]),
    "sagnn_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64p, ctypes.POINTER(vp)]),
    "sagnn_plan_set_interval": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int64, vp]),
    "sagnn_plan_set_latdim_hint": (ctypes.c_int, [vp, ctypes.c_int]),
    "sagnn_plan_set_hot_rows": (ctypes.c_int, [vp, ctypes.c_int]),
    "sagnn_plan_finalize": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "sagnn_plan_get_csr": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "sagnn_plan_get_degrees": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "sagnn_plan_norm_data": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp]),
    "sagnn_plan_get_weights": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp]),
    "sagnn_plan_stats": (ctypes.c_int, [vp, c_i64p]),
    "sagnn_plan_destroy": (ctypes.c_int, [vp]),
    "sagnn_plan_rebalance": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_double), vp]),
    "sagnn_plan_get_split": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_int)]),
    "sagnn_debug_trace": (ctypes.c_int, [vp, vp, ctypes.c_int]),
    "sagnn_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_szp, c_szp, c_szp]),
    "sagnn_propagate_fwd": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           vp, vp, ctypes.c_size_t, vp]),
    "sagnn_propagate_bwd": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           vp, vp, ctypes.c_size_t, vp]),
    "sagnn_propagate_fwd_ex": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                              vp, vp, ctypes.c_size_t, ctypes.c_uint, vp]),
    "sagnn_propagate_bwd_ex": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                              vp, vp, ctypes.c_size_t, ctypes.c_uint, vp]),
    "sagnn_propagate_fwd_scatter": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                   vp, vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                   ctypes.POINTER(vp), ctypes.POINTER(vp), vp]),
    "sagnn_plan_set_row_block": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "sagnn_propagate_fwd_layers": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_float, vp, vp, ctypes.c_size_t, vp]),
    "sagnn_propagate_bwd_levels": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_float, vp, vp, ctypes.c_size_t, vp]),
    "sagnn_workspace_table": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_szp,
                                             c_szp, c_szp]),
    "sagnn_propagate_fwd_interval": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, vp, vp, ctypes.c_size_t, vp]),
    "sagnn_propagate_bwd_interval": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, vp, vp, ctypes.c_size_t, vp]),
    "sagnn_message_propagate": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int,
                                               ctypes.c_float, vp, ctypes.c_size_t, vp]),
    "sagnn_pair_scores_fwd": (ctypes.c_int, [vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_float, vp, vp]),
    "sagnn_pair_scores_bwd": (ctypes.c_int, [vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_float, vp, vp, ctypes.c_int64, vp, ctypes.c_int64, vp]),
    "sagnn_pair_scores_bwd_ws_bytes": (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "sagnn_pair_scores_bwd_det": (ctypes.c_int, [vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_float, vp, vp, ctypes.c_int64, vp, ctypes.c_int64, vp,
                                                 ctypes.c_size_t, vp]),
    "sagnn_sample_ssl_batch": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, vp, vp, vp,
                                              c_i64p, vp]),
    "sagnn_sample_ssl_batch_all": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, vp, vp, vp,
                                                  c_i64p, vp]),
    "sagnn_sample_train_batch": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_uint64, vp, vp, vp, vp, vp, vp, c_i64p, vp]),
    "sagnn_mt19937_seed_numpy": (None, [vp, ctypes.c_uint32]),
    "sagnn_mt19937_seed_python": (None, [vp, ctypes.c_uint64]),
    "sagnn_mt19937_next32": (ctypes.c_uint32, [vp]),
    "sagnn_np_randint": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, vp]),
    "sagnn_np_permutation": (ctypes.c_int, [vp, ctypes.c_int64, vp]),
    "sagnn_py_randint": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int64, c_i64p]),
    "sagnn_np_sample_ssl_batch": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                 vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]),
    "sagnn_np_sample_train_batch": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp,
                                                   vp, vp]),
    "sagnn_host_forward": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int]),
    "sagnn_host_backward": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float]),
    "sagnn_propagate_host": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float]),
}


class SagnnError(RuntimeError):
    """Non-zero status from the C ABI (message from sagnn_last_error)."""

    def __init__(self, code, msg):
        super().__init__("sagnn_b200: %s: %s" % (STATUS.get(code, code), msg))
        self.code = code


def lib_path():
    return _LIB


def load_library():
    """Loads the CUDA library; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise ImportError(
            "sagnn_b200: %s is missing -- build it with `python sa-gnn_b200/build.py` "
            "(nvcc, sm_100a); there is no CPU fallback for the propagation path" % _LIB)
    lib = ctypes.CDLL(_LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load_library().sagnn_last_error()
        raise SagnnError(code, msg.decode() if msg else "")
