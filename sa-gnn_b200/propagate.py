"""Python surface of the propagation path (thin shim over the C ABI).

Keeps the reference's operator surface one level above ``messagePropagate``:

    plan = build_plan(sub_mats, U, I)                      # model.py:227-237 (2T transToLsts calls)
    user_vec, item_vec = propagate(plan, uEmbed, iEmbed,   # model.py:118-129 (T x L x 2 calls,
                                   n_layers=L, leaky=0.5)  #   residuals, add_n)  + autograd backward
    lat = message_propagate(srclats, plan, k, 'user')      # model.py:80-92, one call

All arithmetic happens in ``libsagnn_b200.so`` (hand-written sm_100a kernels); torch only
owns device memory and streams.  No CPU fallback: CPU tensors are rejected.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .data_handler import transToLsts

PAD_ROWS = 100   # model.py:87
_WEIGHT_MODES = {None: 0, "none": 0, "lightgcn": 1, "custom": 2}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Plan:
    """Device-resident CSR (A_k) + transposed CSR (A_k^T) of the T interval graphs plus the
    degree-binned schedule.  Immutable after ``build_plan``; owns no tensors but a scratch cache."""

    def __init__(self, handle, T, U, I, nnz, device, weight_mode, has_val):
        self._h = handle
        self.T, self.U, self.I = T, U, I
        self.nnz = list(nnz)
        self.device = device
        self.weight_mode = weight_mode
        self.has_val = has_val
        self.row_block = None           # (u_begin, u_end, i_begin, i_end) of a row-sharded plan
        self._scratch = {}

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            _lib.load_library().sagnn_plan_destroy(self._h)
            self._h = None
            self._scratch.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("plan already closed")
        return self._h

    # -- parity hooks -----------------------------------------------------------------
    def _rows(self, side):
        return self.I if side else self.U

    def csr(self, k, side):
        """(indptr int32 [R+1], indices int32 [nnz_k]) of A_k (side 0) or A_k^T (side 1)."""
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            indptr = torch.empty(self._rows(side) + 1, dtype=torch.int32, device=self.device)
            indices = torch.empty(self.nnz[k], dtype=torch.int32, device=self.device)
            _lib.check(lib.sagnn_plan_get_csr(self.handle, k, side, _ptr(indptr), _ptr(indices),
                                              _stream_ptr(self.device)))
        return indptr, indices

    def adjacency_list(self, k, side):
        """int32 [nnz_k, 2] (row, col) list, the layout transToLsts returns."""
        indptr, indices = self.csr(k, side)
        counts = (indptr[1:] - indptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(self._rows(side), device=self.device, dtype=torch.int32), counts)
        return torch.stack([rows, indices], dim=1)

    def degrees(self, k, side, value_sum=False):
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            deg = torch.empty(self._rows(side), dtype=torch.int32, device=self.device)
            vs = torch.empty(self._rows(side), dtype=torch.int64, device=self.device) if value_sum else None
            _lib.check(lib.sagnn_plan_get_degrees(self.handle, k, side, _ptr(deg), _ptr(vs),
                                                  _stream_ptr(self.device)))
        return (deg, vs) if value_sum else deg

    def norm_data(self, k, side):
        """The reference's int32-truncated 'normalised' values (DataHandler.py:56-59)."""
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            out = torch.empty(self.nnz[k], dtype=torch.int32, device=self.device)
            _lib.check(lib.sagnn_plan_norm_data(self.handle, k, side, _ptr(out), _stream_ptr(self.device)))
        return out

    def weights(self, k, side):
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            out = torch.empty(self.nnz[k], dtype=torch.float32, device=self.device)
            _lib.check(lib.sagnn_plan_get_weights(self.handle, k, side, _ptr(out), _stream_ptr(self.device)))
        return out

    def stats(self):
        out = (ctypes.c_int64 * 8)()
        _lib.check(_lib.load_library().sagnn_plan_stats(self.handle, out))
        keys = ["rows", "short_rows", "long_rows", "chunks", "max_degree", "edges_both_sides", "hot_rows", "sms"]
        return {k: int(v) for k, v in zip(keys, out) if k != "_"}

    # -- samplers ---------------------------------------------------------------------
    def sample_ssl_batch(self, bat_ids, ssl_num, seed=0):
        """``Recommender.sampleSslBatch(batIds, self.handler.subMat)`` (model.py:304-339) on the device,
        straight from the plan's CSR: returns per interval ``(uLocs, iLocs, uLocs_seq)`` int32 CUDA tensors
        (positives at even, negatives at odd positions; users with fewer than two items in the interval
        emit nothing).  ``seed`` keys the counter-based generator (same seed, same samples)."""
        lib = _lib.load_library()
        bat = _as_dev_i32(bat_ids, self.device)
        batch, cap = int(bat.numel()), max(1, int(bat.numel()) * 2 * int(ssl_num))
        with torch.cuda.device(self.device):
            # all T intervals in one call (one stream synchronisation per step): [T, cap] outputs, counts on the host
            u, i, s = (torch.empty((self.T, cap), dtype=torch.int32, device=self.device) for _ in range(3))
            n = (ctypes.c_int64 * self.T)()
            _lib.check(lib.sagnn_sample_ssl_batch_all(self.handle, _ptr(bat), batch, int(ssl_num), int(seed) & (2**64 - 1),
                                                      _ptr(u), _ptr(i), _ptr(s), n, _stream_ptr(self.device)))
        return [(u[k, :n[k]], i[k, :n[k]], s[k, :n[k]]) for k in range(self.T)]

    def sample_ssl_interval(self, k, bat_ids, ssl_num, seed=0):
        """One interval of ``sample_ssl_batch`` (``sagnn_sample_ssl_batch``): same draws for the same seed."""
        lib = _lib.load_library()
        bat = _as_dev_i32(bat_ids, self.device)
        batch, cap = int(bat.numel()), max(1, int(bat.numel()) * 2 * int(ssl_num))
        with torch.cuda.device(self.device):
            u, i, s = (torch.empty(cap, dtype=torch.int32, device=self.device) for _ in range(3))
            n = ctypes.c_int64()
            _lib.check(lib.sagnn_sample_ssl_batch(self.handle, int(k), _ptr(bat), batch, int(ssl_num), int(seed) & (2**64 - 1),
                                                  _ptr(u), _ptr(i), _ptr(s), ctypes.byref(n), _stream_ptr(self.device)))
        return u[:n.value], i[:n.value], s[:n.value]

    def sample_train_batch(self, bat_ids, sequences, tst_int=None, train_sample_num=40, pred_num=5, pos_length=200,
                           batch_pad=None, seed=0):
        """``Recommender.sampleTrainBatch(batIds, handler.trnMat, ...)`` + ``negSamp`` (model.py:252-302,
        DataHandler.py:28-41) on the device.  ``sequences``: ``handler.sequence`` (list of per-user item lists) or a
        ``(seq_ptr int64 [U+1], seq_items int32)`` pair of tensors; ``tst_int``: ``handler.tstInt`` (None entries = no
        held-out item).  Returns ``(uLocs, iLocs, sequence [batch_pad, pos_length], mask, uLocs_seq, choose)`` CUDA
        tensors laid out like the reference's lists (positives first, then the negatives).  The label test of the
        negative sampler reads the plan's interval CSRs (trnMat = their union); ``seed`` keys the generator."""
        lib = _lib.load_library()
        dev = self.device
        if isinstance(sequences, (tuple, list)) and len(sequences) == 2 and isinstance(sequences[0], torch.Tensor):
            seq_ptr, seq_items = sequences[0].to(dev, torch.int64).contiguous(), sequences[1].to(dev, torch.int32).contiguous()
        else:
            lens = np.fromiter((len(x) for x in sequences), dtype=np.int64, count=len(sequences))
            ptr = np.zeros(len(sequences) + 1, dtype=np.int64)
            np.cumsum(lens, out=ptr[1:])
            flat = np.concatenate([np.asarray(x, dtype=np.int32) for x in sequences]) if ptr[-1] else np.zeros(0, np.int32)
            seq_ptr, seq_items = torch.from_numpy(ptr).to(dev), torch.from_numpy(flat).to(dev)
        if seq_ptr.numel() != self.U + 1:
            raise ValueError("need one sequence per user (%d), got %d" % (self.U, seq_ptr.numel() - 1))
        tst = None
        if tst_int is not None:
            tst = _as_dev_i32(np.array([-1 if x is None else int(x) for x in tst_int], dtype=np.int32), dev)
        bat = _as_dev_i32(bat_ids, dev)
        batch = int(bat.numel())
        batch_pad = batch if batch_pad is None else int(batch_pad)
        cap = max(1, 2 * batch * int(train_sample_num))
        with torch.cuda.device(dev):
            u, i, s = (torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3))
            seq = torch.empty((batch_pad, int(pos_length)), dtype=torch.int32, device=dev)
            mask = torch.empty((batch_pad, int(pos_length)), dtype=torch.float32, device=dev)
            choose = torch.empty(max(batch, 1), dtype=torch.int32, device=dev)
            n = ctypes.c_int64()
            _lib.check(lib.sagnn_sample_train_batch(
                self.handle, _ptr(seq_ptr), _ptr(seq_items), _ptr(tst), _ptr(bat), batch, batch_pad, int(train_sample_num),
                int(pred_num), int(pos_length), int(seed) & (2**64 - 1), _ptr(u), _ptr(i), _ptr(s), _ptr(seq), _ptr(mask),
                _ptr(choose), ctypes.byref(n), _stream_ptr(dev)))
        return u[:n.value], i[:n.value], seq, mask, s[:n.value], choose[:batch]

    # -- scratch ----------------------------------------------------------------------
    def workspace_bytes(self, n_layers, d):
        f, m, b = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.load_library().sagnn_workspace_bytes(self.handle, n_layers, d, ctypes.byref(f),
                                                             ctypes.byref(m), ctypes.byref(b)))
        return f.value, m.value, b.value

    def workspace_table(self, ws, n_layers, d, which, index):
        """(user [T,U,d], item [T,I,d]) float32 views of the workspace table a row-sharded step
        exchanges (``sagnn_workspace_table``): which=0 forward layer output, which=1 backward source."""
        off, ub, ib = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.load_library().sagnn_workspace_table(self.handle, n_layers, d, which, index,
                                                             ctypes.byref(off), ctypes.byref(ub), ctypes.byref(ib)))
        o, nu, ni = off.value, ub.value, ib.value
        return (ws[o:o + nu].view(torch.float32).view(self.T, self.U, d),
                ws[o + nu:o + nu + ni].view(torch.float32).view(self.T, self.I, d))

    def scratch(self, n_layers, d):
        """Cached (workspace tensor, mask_bytes): fwd and bwd of one step run back to back on one
        stream and the backward needs nothing from the forward scratch, so they share it.  The workspace
        holds tickets, queue heads, slice partials and the ping-pong tables of a call in flight, so it is
        cached PER CUDA STREAM: calls issued on different streams never share one."""
        key = (n_layers, d, torch.cuda.current_stream(self.device).cuda_stream)
        if key not in self._scratch:
            f, m, b = self.workspace_bytes(n_layers, d)
            ws = torch.empty(max(f, b, 1), dtype=torch.uint8, device=self.device)
            self._scratch[key] = (ws, m)
        return self._scratch[key]


def _as_dev_i32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(device)


def _split_interval(m):
    """-> (row, col, val|None, shape|None) for one interval given in any accepted form."""
    if hasattr(m, "tocoo"):                       # scipy sparse: the reference's own path
        idx, data, shape = transToLsts(m)         # includes the (0,0) fallback for empty matrices
        return idx[:, 0], idx[:, 1], data, shape
    if isinstance(m, (tuple, list)):
        if len(m) not in (2, 3):
            raise ValueError("interval tuple must be (row, col) or (row, col, val)")
        if len(m[0]) == 0:                                 # empty interval: the fallback edge of DataHandler.py:66-68
            z = np.zeros(1, dtype=np.int32)
            return z, z, (z if len(m) == 3 else None), None
        return m[0], m[1], (m[2] if len(m) == 3 else None), None
    if getattr(m, "ndim", 0) == 2 and m.shape[1] == 2:     # transToLsts-style [E,2] list
        if m.shape[0] == 0:
            m = np.array([[0, 0]], dtype=np.int32)         # DataHandler.py:66-68
        return m[:, 0], m[:, 1], None, None
    raise TypeError("unsupported interval description: %r" % type(m))


def build_plan(sub_mats, U=None, I=None, *, device=None, edge_weight=None, strict_pad=False, latdim=64,
               row_block=None, padded_shape=None, hot_rows=0):
    """Builds the device plan for ``T = len(sub_mats)`` interval graphs.

    sub_mats[k]: scipy sparse ``U x I`` matrix (``handler.subMat[k]``), or an ``[E,2]`` adjacency
    list as ``transToLsts`` returns it, or ``(row, col[, val])`` arrays / tensors, row-major sorted.
    edge_weight: None (reference-exact binary structure), ``"lightgcn"`` (1/sqrt(d_u d_i)) or a
    list of per-interval fp32 arrays in the adjacency list's order.
    latdim: the embedding width the plan will mostly run with (picks the kernel and its schedule).
    hot_rows: stage that many highest-degree rows of every source table in shared memory
    (``sagnn_plan_set_hot_rows``; default 0 -- measured slower on B200).
    strict_pad: raise like TF-CPU does when an interval's last populated row is more than 100
    rows before the end (model.py:87-91); by default such rows are simply zero (TF-GPU).
    row_block: ``(u_begin, u_end, i_begin, i_end)`` -- row sharding (``sagnn_plan_set_row_block``): the
    plan only computes those user / item rows of every interval (see ``dist.RowShardedPropagation``).
    padded_shape: ``(U_pad, I_pad) >= (U, I)``: tables get extra empty rows (equal row blocks per rank).
    """
    if not torch.cuda.is_available():
        raise RuntimeError("sagnn_b200.build_plan needs a CUDA device (there is no CPU fallback)")
    lib = _lib.load_library()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    T = len(sub_mats)
    if T == 0:
        raise ValueError("need at least one interval graph")
    parts = [_split_interval(m) for m in sub_mats]
    for _, _, _, shape in parts:
        if shape is not None:
            U = shape[0] if U is None else U
            I = shape[1] if I is None else I
            if (U, I) != (shape[0], shape[1]):
                raise ValueError("interval shape %s does not match (U, I) = (%d, %d)" % (shape, U, I))
    if U is None or I is None:
        raise ValueError("U and I are required when intervals are given as index arrays")
    if padded_shape is not None:
        if padded_shape[0] < U or padded_shape[1] < I:
            raise ValueError("padded_shape %s is smaller than (U, I) = (%d, %d)" % (tuple(padded_shape), U, I))
        U, I = int(padded_shape[0]), int(padded_shape[1])
    has_val = all(p[2] is not None for p in parts)
    custom = isinstance(edge_weight, (list, tuple))
    mode = 2 if custom else _WEIGHT_MODES[edge_weight]
    if custom and len(edge_weight) != T:
        raise ValueError("need one weight array per interval")
    nnz = [int(len(p[0])) for p in parts]
    if strict_pad:
        for k, (row, col, _, _) in enumerate(parts):
            rmax = int(row[-1]) if nnz[k] else 0
            cmax = int(col.max()) if nnz[k] else 0
            if rmax + 1 + PAD_ROWS < U or cmax + 1 + PAD_ROWS < I:
                raise IndexError("interval %d: padded segment_sum has fewer rows than the lookup "
                                 "range (model.py:87-91 would raise on TF-CPU)" % k)
    handle = ctypes.c_void_p()
    with torch.cuda.device(device):
        nnz_arr = (ctypes.c_int64 * T)(*nnz)
        _lib.check(lib.sagnn_plan_create(T, int(U), int(I), nnz_arr, ctypes.byref(handle)))
        plan = Plan(handle, T, int(U), int(I), nnz, device, mode, has_val)
        st = _stream_ptr(device)
        for k, (row, col, val, _) in enumerate(parts):
            row_d, col_d = _as_dev_i32(row, device), _as_dev_i32(col, device)
            val_d = _as_dev_i32(val, device) if has_val else None
            w_d = None
            if custom:
                w = edge_weight[k]
                w_d = (w.to(device=device, dtype=torch.float32) if isinstance(w, torch.Tensor)
                       else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).to(device)).contiguous()
                if w_d.numel() != nnz[k]:
                    raise ValueError("interval %d: %d weights for %d edges" % (k, w_d.numel(), nnz[k]))
            _lib.check(lib.sagnn_plan_set_interval(handle, k, _ptr(row_d), _ptr(col_d), _ptr(val_d),
                                                   _ptr(w_d), nnz[k], st))
        if row_block is not None:
            _lib.check(lib.sagnn_plan_set_row_block(handle, *[int(x) for x in row_block]))
            plan.row_block = tuple(int(x) for x in row_block)
        _lib.check(lib.sagnn_plan_set_latdim_hint(handle, int(latdim)))
        if hot_rows:
            _lib.check(lib.sagnn_plan_set_hot_rows(handle, int(hot_rows)))
        _lib.check(lib.sagnn_plan_finalize(handle, mode, st))
    return plan


def bucket_events(users, items, times, U, I, graph_num, minn=None, maxx=None, device=None):
    """Device-side ``trans_sub`` (preprocess_to_trnmat.ipynb cell 7; ``sagnn_bucket_events``): raw
    ``(user, item, timestamp)`` events in the notebook's visiting order -> the ``graph_num`` interval
    adjacency lists as ``[(row, col, val)]`` int32 CUDA tensors (row-major sorted, first-event
    timestamps), ready for ``build_plan(lists, U, I)``.  ``minn`` / ``maxx`` default to what `trans`
    (cell 13) computes.  Same result as ``data_handler.trans_sub``, bit for bit."""
    if not torch.cuda.is_available():
        raise RuntimeError("sagnn_b200.bucket_events needs a CUDA device (there is no CPU fallback)")
    from .data_handler import TS_MAXX_INIT, TS_MINN_INIT
    lib = _lib.load_library()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = (times if isinstance(times, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(times, dtype=np.int64)))
    t = t.to(device=device, dtype=torch.int64).contiguous()
    u, i = _as_dev_i32(users, device), _as_dev_i32(items, device)
    n = int(t.numel())
    if u.numel() != n or i.numel() != n:
        raise ValueError("users, items and times must have the same length")
    if minn is None:
        minn = min(TS_MINN_INIT, int(t.min())) if n else TS_MINN_INIT
    if maxx is None:
        maxx = max(TS_MAXX_INIT, int(t.max())) if n else TS_MAXX_INIT
    T = int(graph_num)
    with torch.cuda.device(device):
        row = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        col, val = torch.empty_like(row), torch.empty_like(row)
        nnz = (ctypes.c_int64 * T)()
        _lib.check(lib.sagnn_bucket_events(_ptr(u), _ptr(i), _ptr(t), n, int(U), int(I), T, int(minn), int(maxx),
                                           _ptr(row), _ptr(col), _ptr(val), nnz, _stream_ptr(device)))
    out, off = [], 0
    for k in range(T):
        out.append((row[off:off + nnz[k]], col[off:off + nnz[k]], val[off:off + nnz[k]]))
        off += nnz[k]
    return out


def _check_tables(plan, u, i):
    if not (u.is_cuda and i.is_cuda):
        raise RuntimeError("sagnn_b200.propagate: embeddings must be CUDA tensors (no CPU fallback)")
    if u.dtype != torch.float32 or i.dtype != torch.float32:
        raise TypeError("embeddings must be float32")
    if u.dim() != 3 or i.dim() != 3 or u.shape[0] != plan.T or i.shape[0] != plan.T or \
            u.shape[1] != plan.U or i.shape[1] != plan.I or u.shape[2] != i.shape[2]:
        raise ValueError("expected uEmbed [%d,%d,d] and iEmbed [%d,%d,d], got %s and %s"
                         % (plan.T, plan.U, plan.T, plan.I, tuple(u.shape), tuple(i.shape)))
    if u.device != plan.device or i.device != plan.device:
        raise ValueError("embeddings live on %s but the plan on %s" % (u.device, plan.device))


_LAYOUTS = {"trd": 0, "rtd": 1}   # SAGNN_LAYOUT_TRD / SAGNN_LAYOUT_RTD


class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_embed, i_embed, plan, n_layers, leaky, layout):
        _check_tables(plan, u_embed, i_embed)
        lib = _lib.load_library()
        u, i = u_embed.contiguous(), i_embed.contiguous()
        d = u.shape[2]
        need_bwd = u_embed.requires_grad or i_embed.requires_grad
        with torch.cuda.device(plan.device):
            ws, mask_bytes = plan.scratch(n_layers, d)
            if layout:
                u_out = torch.empty((plan.U, plan.T, d), dtype=torch.float32, device=plan.device)
                i_out = torch.empty((plan.I, plan.T, d), dtype=torch.float32, device=plan.device)
            else:
                u_out, i_out = torch.empty_like(u), torch.empty_like(i)
            masks = torch.empty(mask_bytes, dtype=torch.uint8, device=plan.device) if need_bwd else None
            _lib.check(lib.sagnn_propagate_fwd_ex(plan.handle, _ptr(u), _ptr(i), _ptr(u_out), _ptr(i_out),
                                                  n_layers, d, float(leaky), _ptr(masks), _ptr(ws), ws.numel(),
                                                  layout, _stream_ptr(plan.device)))
        ctx.plan, ctx.n_layers, ctx.leaky, ctx.d, ctx.layout = plan, n_layers, float(leaky), d, layout
        ctx.masks = masks
        return u_out, i_out

    @staticmethod
    def backward(ctx, g_user, g_item):
        plan, d = ctx.plan, ctx.d
        lib = _lib.load_library()
        su, si = ((plan.U, plan.T, d), (plan.I, plan.T, d)) if ctx.layout else ((plan.T, plan.U, d), (plan.T, plan.I, d))
        with torch.cuda.device(plan.device):
            if g_user is None:
                g_user = torch.zeros(su, dtype=torch.float32, device=plan.device)
            if g_item is None:
                g_item = torch.zeros(si, dtype=torch.float32, device=plan.device)
            g_user, g_item = g_user.contiguous(), g_item.contiguous()
            ws, _ = plan.scratch(ctx.n_layers, d)
            d_u = torch.empty((plan.T, plan.U, d), dtype=torch.float32, device=plan.device)
            d_i = torch.empty((plan.T, plan.I, d), dtype=torch.float32, device=plan.device)
            _lib.check(lib.sagnn_propagate_bwd_ex(plan.handle, _ptr(g_user), _ptr(g_item), _ptr(d_u), _ptr(d_i),
                                                  ctx.n_layers, d, ctx.leaky, _ptr(ctx.masks), _ptr(ws),
                                                  ws.numel(), ctx.layout, _stream_ptr(plan.device)))
        return d_u, d_i, None, None, None, None


def propagate(plan, u_embed, i_embed, n_layers, leaky=0.5, layout="trd"):
    """model.py:118-129 for all T intervals: returns (user_vector [T,U,d], item_vector [T,I,d]),
    i.e. the stacked ``tf.add_n(embs0)`` / ``tf.add_n(embs1)`` of model.py:126-132.  Differentiable
    w.r.t. both embedding tables (dense upstream gradients).  Edge dropout
    (model.py:93-102) only rewrites the ignored edge values, so it has no counterpart here.

    layout="rtd" returns ``user_vector_tensor [U,T,d]`` / ``item_vector_tensor [I,T,d]`` instead,
    the ``tf.transpose(.., perm=[1,0,2])`` of model.py:133-134 that feeds the LSTM: the transpose is
    fused into the kernel epilogue (and the backward reads its upstream in that layout), values are
    bitwise those of the default layout."""
    if layout not in _LAYOUTS:
        raise ValueError("layout must be 'trd' ([T,R,d], model.py:131-132) or 'rtd' ([R,T,d], model.py:133-134)")
    return _Propagate.apply(u_embed, i_embed, plan, int(n_layers), float(leaky), _LAYOUTS[layout])


def message_propagate(srclats, plan, k, type="user", leaky=0.5):
    """One ``messagePropagate(srclats, mat, type)`` call (model.py:80-92) on interval ``k``:
    ``type='user'`` uses subAdj[k] (srclats = item table [I,d] -> [U,d]), ``type='item'`` uses
    subTpAdj[k] (srclats = user table [U,d] -> [I,d]).  Forward only (op-level parity hook)."""
    if not srclats.is_cuda:
        raise RuntimeError("sagnn_b200.message_propagate: srclats must be a CUDA tensor")
    side = 0 if type == "user" else 1
    rows, src_rows = (plan.U, plan.I) if side == 0 else (plan.I, plan.U)
    if srclats.dim() != 2 or srclats.shape[0] != src_rows or srclats.dtype != torch.float32:
        raise ValueError("srclats must be float32 [%d, d]" % src_rows)
    lib = _lib.load_library()
    src = srclats.contiguous()
    d = src.shape[1]
    with torch.cuda.device(plan.device):
        ws, _ = plan.scratch(1, d)
        out = torch.empty((rows, d), dtype=torch.float32, device=plan.device)
        _lib.check(lib.sagnn_message_propagate(plan.handle, int(k), side, _ptr(src), _ptr(out), d, float(leaky),
                                               _ptr(ws), ws.numel(), _stream_ptr(plan.device)))
    return out


class _PairScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, user_vec, item_vec, k, uids, iids, activation, leaky, layout, deterministic=True):
        lib = _lib.load_library()
        u, i = user_vec.contiguous(), item_vec.contiguous()
        d = u.shape[2]
        n = int(uids.numel())
        scores = torch.empty(n, dtype=torch.float32, device=u.device)
        geo = _pair_geometry(u, i, k, layout)
        with torch.cuda.device(u.device):
            _lib.check(lib.sagnn_pair_scores_fwd(geo[0], geo[1], geo[2], geo[3], _ptr(uids), _ptr(iids), n, d,
                                                 activation, leaky, _ptr(scores), _stream_ptr(u.device)))
        ctx.save_for_backward(u, i, uids, iids)
        ctx.args = (k, activation, leaky, layout, n, d, deterministic)
        return scores

    @staticmethod
    def backward(ctx, g):
        u, i, uids, iids = ctx.saved_tensors
        k, activation, leaky, layout, n, d, deterministic = ctx.args
        lib = _lib.load_library()
        d_u = torch.zeros_like(u) if ctx.needs_input_grad[0] else None
        d_i = torch.zeros_like(i) if ctx.needs_input_grad[1] else None
        su = _pair_geometry(u, i, k, layout)                                  # the tables
        gu = _pair_geometry(d_u if d_u is not None else u, d_i if d_i is not None else i, k, layout)   # their gradients
        with torch.cuda.device(u.device):
            if deterministic:     # sorted by destination row, one writer per row: no float atomics, same bits every run
                nb = ctypes.c_size_t()
                _lib.check(lib.sagnn_pair_scores_bwd_ws_bytes(n, ctypes.byref(nb)))
                ws = torch.empty(max(1, nb.value), dtype=torch.uint8, device=u.device)
                _lib.check(lib.sagnn_pair_scores_bwd_det(su[0], su[1], su[2], su[3], _ptr(uids), _ptr(iids), n, d, activation,
                                                         leaky, _ptr(g.contiguous()), gu[0] if d_u is not None else None, gu[1],
                                                         gu[2] if d_i is not None else None, gu[3], _ptr(ws), nb.value,
                                                         _stream_ptr(u.device)))
            else:
                _lib.check(lib.sagnn_pair_scores_bwd(su[0], su[1], su[2], su[3], _ptr(uids), _ptr(iids), n, d, activation,
                                                     leaky, _ptr(g.contiguous()), gu[0] if d_u is not None else None, gu[1],
                                                     gu[2] if d_i is not None else None, gu[3], _stream_ptr(u.device)))
        return d_u, d_i, None, None, None, None, None, None, None


def _pair_geometry(u, i, k, layout):
    """(user base, user row stride, item base, item row stride) of interval k: [T,R,d] (layout 0) or [R,T,d]."""
    d = u.shape[2]
    if layout:      # [R,T,d]: interval k starts k*d floats in, rows are T*d apart
        T = u.shape[1]
        return (ctypes.c_void_p(u.data_ptr() + 4 * k * d), T * d, ctypes.c_void_p(i.data_ptr() + 4 * k * d), T * d)
    return (ctypes.c_void_p(u.data_ptr() + 4 * k * u.shape[1] * d), d,
            ctypes.c_void_p(i.data_ptr() + 4 * k * i.shape[1] * d), d)


def pair_scores(user_vec, item_vec, k, uids, iids, activation="leakyRelu", leaky=0.5, layout="trd", check_ids=True,
                deterministic=True):
    """``sum_c act(user_vec[k][uids] * item_vec[k][iids])`` over the propagation's outputs: the SSL
    scores ``preds_one`` of interval ``k`` (model.py:194-198; ``activation="leakyRelu"``) or a plain
    prediction dot product (model.py:171-173; ``activation=None``).  ``user_vec`` / ``item_vec`` are
    ``[T,U,d]`` / ``[T,I,d]`` (``layout="trd"``) or ``[U,T,d]`` / ``[I,T,d]`` (``"rtd"``); ``uids`` / ``iids`` int
    CUDA tensors of equal length (``suids[k]`` / ``siids[k]``).  Differentiable w.r.t. both tables: the
    backward scatters the sparse gradient with a hand-written kernel: by default the atomic-free one
    (``sagnn_pair_scores_bwd_det``: samples sorted by destination row, one writer per row, same bits every run);
    ``deterministic=False`` uses float atomics (``sagnn_pair_scores_bwd``)."""
    if not (user_vec.is_cuda and item_vec.is_cuda and uids.is_cuda and iids.is_cuda):
        raise RuntimeError("sagnn_b200.pair_scores: tensors must be CUDA tensors (no CPU fallback)")
    if user_vec.dtype != torch.float32 or item_vec.dtype != torch.float32:
        raise TypeError("tables must be float32")
    if layout not in _LAYOUTS or activation not in (None, "none", "leakyRelu"):
        raise ValueError("layout must be 'trd' or 'rtd', activation None or 'leakyRelu'")
    lay = _LAYOUTS[layout]
    T, U, I = (user_vec.shape[1], user_vec.shape[0], item_vec.shape[0]) if lay else \
        (user_vec.shape[0], user_vec.shape[1], item_vec.shape[1])
    if not 0 <= int(k) < T or uids.numel() != iids.numel() or user_vec.shape[2] != item_vec.shape[2]:
        raise ValueError("bad interval / id arrays / latdim")
    uids, iids = uids.to(torch.int32).contiguous(), iids.to(torch.int32).contiguous()
    if check_ids and uids.numel():
        if int(uids.min()) < 0 or int(uids.max()) >= U or int(iids.min()) < 0 or int(iids.max()) >= I:
            raise IndexError("pair_scores: id out of range (tf.nn.embedding_lookup on CPU raises too)")
    act = 1 if activation == "leakyRelu" else 0
    return _PairScores.apply(user_vec, item_vec, int(k), uids, iids, act, float(leaky), lay, bool(deterministic))


def _host_ptr(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous():
            raise ValueError("host buffers must be contiguous float32 CPU tensors")
        return ctypes.c_void_p(a.data_ptr())
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("host buffers must be C-contiguous float32 arrays")
    return ctypes.c_void_p(a.ctypes.data)


def propagate_host(plan, u_embed, i_embed, g_user, g_item, user_out, item_out, d_u, d_i, n_layers, leaky=0.5):
    """Host-buffer entry point (``sagnn_propagate_host``): numpy arrays or (pinned) CPU tensors in,
    results written into the given host buffers; H2D/D2H copies happen inside the call.
    ``g_user``/``g_item``/``d_u``/``d_i`` may be None for forward only."""
    d = int(u_embed.shape[2])
    with torch.cuda.device(plan.device):
        _lib.check(_lib.load_library().sagnn_propagate_host(
            plan.handle, _host_ptr(u_embed), _host_ptr(i_embed), _host_ptr(g_user), _host_ptr(g_item),
            _host_ptr(user_out), _host_ptr(item_out), _host_ptr(d_u), _host_ptr(d_i), int(n_layers), d,
            float(leaky)))


def host_forward(plan, u_embed, i_embed, user_out, item_out, n_layers, leaky=0.5, keep_masks=True):
    """``sagnn_host_forward``: host buffers in/out, forward only; the sign masks stay inside the plan
    for a later ``host_backward`` (what a TF1 ``py_func`` forward op would call)."""
    with torch.cuda.device(plan.device):
        _lib.check(_lib.load_library().sagnn_host_forward(
            plan.handle, _host_ptr(u_embed), _host_ptr(i_embed), _host_ptr(user_out), _host_ptr(item_out),
            int(n_layers), int(u_embed.shape[2]), float(leaky), 1 if keep_masks else 0))


def host_backward(plan, g_user, g_item, d_u, d_i, n_layers, leaky=0.5):
    """``sagnn_host_backward``: gradients for the last ``host_forward(keep_masks=True)`` on this plan."""
    with torch.cuda.device(plan.device):
        _lib.check(_lib.load_library().sagnn_host_backward(
            plan.handle, _host_ptr(g_user), _host_ptr(g_item), _host_ptr(d_u), _host_ptr(d_i),
            int(n_layers), int(g_user.shape[2]), float(leaky)))


class IntervalPropagation(torch.nn.Module):
    """The short-term graph-propagation block of ``Recommender.ours()`` (model.py:108-134) as a
    module: owns ``uEmbed [T,U,d]`` / ``iEmbed [T,I,d]`` (xavier, model.py:108-109) and returns the
    stacked per-interval layer sums."""

    def __init__(self, plan, latdim=64, gnn_layer=2, leaky=0.5, layout="trd"):
        super().__init__()
        self.plan, self.gnn_layer, self.leaky, self.layout = plan, gnn_layer, leaky, layout
        self.uEmbed = torch.nn.Parameter(torch.empty(plan.T, plan.U, latdim, device=plan.device))
        self.iEmbed = torch.nn.Parameter(torch.empty(plan.T, plan.I, latdim, device=plan.device))
        # tf.contrib xavier on a 3-D shape: fan_in = T*rows, fan_out = T*d (Utils/NNLayers.py:47-50)
        for p_ in (self.uEmbed, self.iEmbed):
            a = (6.0 / (plan.T * (p_.shape[1] + latdim))) ** 0.5
            torch.nn.init.uniform_(p_, -a, a)

    def forward(self):
        return propagate(self.plan, self.uEmbed, self.iEmbed, self.gnn_layer, self.leaky, self.layout)
