"""ctypes loader for the C restatement (TEST INFRASTRUCTURE ONLY; see csrc/sagnn_oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsagnn_oracle.so")
_lib = None


def build(force=False):
    """gcc the C oracle in place (used by __graft_entry__.build())."""
    src = os.path.join(_HERE, "csrc", "sagnn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsagnn_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.sagnn_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().sagnn_oracle_num_threads())


def set_threads(n):
    lib().sagnn_oracle_set_threads(ctypes.c_int(int(n)))


def indices_to_csr(indices, n_rows):
    """row-major sorted adjacency list [E,2] -> (indptr int64 [R+1], idx int32 [E])."""
    indices = np.asarray(indices)
    counts = np.bincount(indices[:, 0], minlength=n_rows)
    ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr, np.ascontiguousarray(indices[:, 1], dtype=np.int32)


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def propagate(adj, tp_adj, u_embed, i_embed, g_user, g_item, n_layers, leaky=0.5, dtype=np.float64,
              edge_weight=None, tp_edge_weight=None, mask_in=None, mask_cmp=None, tie_tol=1e-5,
              return_masks=False):
    """Same contract as propagate_oracle.propagate; g_user/g_item may be None (forward only).
    Returns (user_vec, item_vec, dU, dI) (dU/dI None when forward only).

    Masks are uint8 arrays [T, L, (U+I)*d] (user part then item part; 1 = gradient passes
    unscaled).  mask_in overrides the oracle's own MaximumGrad decisions in the backward;
    mask_cmp is compared against them and the result carries ``stats = [mismatches,
    mismatches that are not near-ties (|z| > tie_tol * sum|terms|)]``; return_masks appends
    the oracle's masks.  Extra results are returned as a dict in a 5th tuple slot."""
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        fn, ct = lib().sagnn_oracle_interval_f64, ctypes.c_double
    elif dtype == np.float32:
        fn, ct = lib().sagnn_oracle_interval_f32, ctypes.c_float
    else:
        raise TypeError(dtype)
    fn.restype = ctypes.c_int
    T, U, d = u_embed.shape
    I = i_embed.shape[1]
    uE = np.ascontiguousarray(u_embed, dtype=dtype)
    iE = np.ascontiguousarray(i_embed, dtype=dtype)
    bwd = g_user is not None
    gU = np.ascontiguousarray(g_user, dtype=dtype) if bwd else None
    gI = np.ascontiguousarray(g_item, dtype=dtype) if bwd else None
    uO, iO = np.empty_like(uE), np.empty_like(iE)
    dU = np.empty_like(uE) if bwd else None
    dI = np.empty_like(iE) if bwd else None
    extra = mask_in is not None or mask_cmp is not None or return_masks
    m_out = np.zeros((T, n_layers, (U + I) * d), dtype=np.uint8) if return_masks else None
    stats = np.zeros(2, dtype=np.int64)
    u8 = ctypes.c_uint8
    for k in range(T):
        uptr, ucol = indices_to_csr(adj[k], U)
        iptr, irow = indices_to_csr(tp_adj[k], I)
        uw = None if edge_weight is None else np.ascontiguousarray(edge_weight[k], dtype=dtype)
        iw = None if tp_edge_weight is None else np.ascontiguousarray(tp_edge_weight[k], dtype=dtype)
        rc = fn(ctypes.c_int(U), ctypes.c_int(I), ctypes.c_int(d), ctypes.c_int(n_layers),
                ctypes.c_double(leaky),
                _p(uptr, ctypes.c_int64), _p(ucol, ctypes.c_int32),
                _p(iptr, ctypes.c_int64), _p(irow, ctypes.c_int32),
                _p(uw, ct), _p(iw, ct),
                _p(uE[k], ct), _p(iE[k], ct),
                _p(gU[k], ct) if bwd else None, _p(gI[k], ct) if bwd else None,
                _p(uO[k], ct), _p(iO[k], ct),
                _p(dU[k], ct) if bwd else None, _p(dI[k], ct) if bwd else None,
                _p(m_out[k], u8) if m_out is not None else None,
                _p(np.ascontiguousarray(mask_in[k], dtype=np.uint8), u8) if mask_in is not None else None,
                _p(np.ascontiguousarray(mask_cmp[k], dtype=np.uint8), u8) if mask_cmp is not None else None,
                ctypes.c_double(tie_tol), _p(stats, ctypes.c_int64))
        if rc != 0:
            raise MemoryError("sagnn_oracle_interval failed")
    if extra:
        return uO, iO, dU, dI, {"stats": stats, "masks": m_out}
    return uO, iO, dU, dI
