"""Stream-exact samplers (parity mode of SURVEY 8f N3): ``Recommender.sampleSslBatch`` / ``sampleTrainBatch`` /
``negSamp`` (LIU-YUXI/SA-GNN ``model.py:252-339``, ``DataHandler.py:28-41``) with the reference's own random draws.

The reference samples from numpy's global ``RandomState`` and CPython's ``random`` module, both seeded in
``main.py:21-22``.  ``ReferenceStream`` holds one MT19937 state for each, seeded the same way (or taken over from
``np.random.get_state()`` / ``random.getstate()``), and the C ABI functions ``sagnn_np_sample_ssl_batch`` /
``sagnn_np_sample_train_batch`` (``csrc/np_stream.cu``, host code: the streams are sequential) walk the caller's
scipy CSR arrays instead of densifying ``labelMat[batIds].toarray()``.  Same seeds, same samples, same stream
positions afterwards -- so a reference run can switch samplers mid-epoch (``to_numpy()`` / ``to_python()`` hand the
states back).  The device samplers (``Plan.sample_ssl_batch`` / ``Plan.sample_train_batch``) are the throughput mode:
same output contract, their own counter-based stream.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class _MT(ctypes.Structure):
    _fields_ = [("key", ctypes.c_uint32 * 624), ("pos", ctypes.c_int32)]


def _csr_arrays(m):
    """(indptr int32, indices int32, nonzero uint8 | None) of a scipy CSR matrix in canonical format."""
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
    indptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(m.indices, dtype=np.int32)
    nz = None
    if m.nnz and not np.all(m.data != 0):          # stored zeros: `temLabel != 0` / `temLabel[item] == 0` see through them
        nz = np.ascontiguousarray(m.data != 0, dtype=np.uint8)
    return indptr, indices, nz


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


class ReferenceStream:
    """The two generator states of a reference run: ``np`` (numpy's global RandomState) and ``py`` (``random``)."""

    def __init__(self, np_seed=100, py_seed=100):
        self._lib = _lib.load_library()
        self._np, self._py = _MT(), _MT()
        self._lib.sagnn_mt19937_seed_numpy(ctypes.byref(self._np), int(np_seed))
        self._lib.sagnn_mt19937_seed_python(ctypes.byref(self._py), int(py_seed))

    # -- hand-over with the interpreters' own generators ---------------------------------------
    def from_numpy(self, state=None):
        """Takes over ``np.random.get_state()`` (default: the current global state)."""
        st = np.random.get_state() if state is None else state
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        ctypes.memmove(self._np.key, key.ctypes.data, 624 * 4)
        self._np.pos = int(st[2])
        return self

    def to_numpy(self):
        """A tuple for ``np.random.set_state`` (no cached Gaussian: none of the sampler's calls makes one)."""
        return ("MT19937", np.frombuffer(self._np.key, dtype=np.uint32).copy(), int(self._np.pos), 0, 0.0)

    def from_python(self, state=None):
        import random
        st = random.getstate() if state is None else state
        key = np.asarray(st[1][:624], dtype=np.uint32)
        ctypes.memmove(self._py.key, key.ctypes.data, 624 * 4)
        self._py.pos = int(st[1][624])
        return self

    def to_python(self):
        """A tuple for ``random.setstate``."""
        return (3, tuple(int(x) for x in np.frombuffer(self._py.key, dtype=np.uint32)) + (int(self._py.pos),), None)

    # -- primitives --------------------------------------------------------------------------
    def np_randint(self, low, high, size=None):
        """``np.random.randint(low, high, size)`` (``np.random.choice(n)`` is ``np_randint(0, n)``)."""
        n = 1 if size is None else int(size)
        out = np.empty(n, dtype=np.int64)
        _lib.check(self._lib.sagnn_np_randint(ctypes.byref(self._np), int(low), int(high), n, _p(out)))
        return int(out[0]) if size is None else out

    def np_permutation(self, n):
        out = np.empty(int(n), dtype=np.int64)
        _lib.check(self._lib.sagnn_np_permutation(ctypes.byref(self._np), int(n), _p(out)))
        return out

    def py_randint(self, a, b):
        out = ctypes.c_int64()
        _lib.check(self._lib.sagnn_py_randint(ctypes.byref(self._py), int(a), int(b), ctypes.byref(out)))
        return out.value

    # -- static inputs are converted once (handler.subMat / trnMat / sequence / tstInt do not change in a run) ------
    def _cached(self, obj, make):
        if not hasattr(self, "_cache"):
            self._cache = {}
        hit = self._cache.get(id(obj))
        if hit is None or hit[0] is not obj:
            hit = (obj, make(obj))            # keeps obj alive, so its id cannot be reused
            self._cache[id(obj)] = hit
        return hit[1]

    @staticmethod
    def _seq_csr(sequences):
        lens = np.fromiter((len(x) for x in sequences), dtype=np.int64, count=len(sequences))
        ptr = np.zeros(len(sequences) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        flat = np.concatenate([np.asarray(x, dtype=np.int32) for x in sequences]) if ptr[-1] else np.zeros(0, np.int32)
        return ptr, flat

    # -- the samplers ------------------------------------------------------------------------
    def sample_ssl_batch(self, bat_ids, sub_mats, ssl_num, n_item=None):
        """``sampleSslBatch(batIds, handler.subMat)`` with ``args.sslNum = ssl_num``, ``args.item = n_item``:
        returns ``(uLocs, iLocs, uLocs_seq)``, each a list of T int32 arrays (the reference's lists of lists)."""
        T = len(sub_mats)
        U, I = sub_mats[0].shape
        n_item = I if n_item is None else int(n_item)
        csr = [self._cached(m, _csr_arrays) for m in sub_mats]
        arr = lambda j: (ctypes.c_void_p * T)(*[_p(c[j]) for c in csr])
        bat = np.ascontiguousarray(bat_ids, dtype=np.int32)
        u, i, s = (np.empty((T, len(bat) * 2 * int(ssl_num)), dtype=np.int32) for _ in range(3))
        n = np.zeros(T, dtype=np.int64)
        nzs = arr(2) if any(c[2] is not None for c in csr) else None
        _lib.check(self._lib.sagnn_np_sample_ssl_batch(ctypes.byref(self._np), T, arr(0), arr(1), nzs, _p(bat), len(bat),
                                                       int(ssl_num), U, n_item, _p(u), _p(i), _p(s), _p(n)))
        cut = lambda a: [a[k, :n[k]].copy() for k in range(T)]
        return cut(u), cut(i), cut(s)

    def sample_train_batch(self, bat_ids, label_mat, sequences, tst_int, train_sample_num, pred_num=5, pos_length=200,
                           batch_pad=None, n_item=None):
        """``sampleTrainBatch(batIds, handler.trnMat, handler.timeMat, train_sample_num)`` with ``args.pred_num``,
        ``args.pos_length``, ``args.batch = batch_pad``, ``args.item = n_item``; ``sequences`` = ``handler.sequence``,
        ``tst_int`` = ``handler.tstInt`` (None entries allowed).  Returns ``(uLocs, iLocs, sequence, mask, uLocs_seq)``
        like the reference (int32 index arrays; ``sequence`` int64 and ``mask`` float64 ``[batch_pad, pos_length]``).
        The static inputs (matrices, sequences, tst_int) are converted to flat arrays on first use and cached by
        object identity: pass the same objects every step, as the reference's handler does."""
        U, I = label_mat.shape
        n_item = I if n_item is None else int(n_item)
        ptr, flat = self._cached(sequences, self._seq_csr)
        tst = None if tst_int is None else self._cached(
            tst_int, lambda t: np.array([-1 if x is None else int(x) for x in t], dtype=np.int32))
        indptr, indices, nz = self._cached(label_mat, _csr_arrays)
        bat = np.ascontiguousarray(bat_ids, dtype=np.int32)
        batch = len(bat)
        batch_pad = batch if batch_pad is None else int(batch_pad)
        cap = 2 * batch * int(train_sample_num)
        u, i, s = (np.empty(cap, dtype=np.int32) for _ in range(3))
        seq = np.empty((batch_pad, int(pos_length)), dtype=np.int64)
        mask = np.empty((batch_pad, int(pos_length)), dtype=np.float64)
        n = ctypes.c_int64()
        _lib.check(self._lib.sagnn_np_sample_train_batch(
            ctypes.byref(self._np), ctypes.byref(self._py), _p(ptr), _p(flat), _p(tst), _p(indptr), _p(indices), _p(nz),
            _p(bat), batch, batch_pad, int(train_sample_num), int(pred_num), int(pos_length), len(sequences), n_item,
            _p(u), _p(i), _p(s), _p(seq), _p(mask), None, ctypes.byref(n)))
        return u[:n.value].copy(), i[:n.value].copy(), seq, mask, s[:n.value].copy()
