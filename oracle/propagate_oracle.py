"""numpy restatement of the reference's interval-graph propagation (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to the
upstream checkout, LIU-YUXI/SA-GNN).  The restatement is written op by op in
the order TensorFlow 1.14 executes the graph that ``model.py`` builds, so it is
deliberately *un*-fused: it materialises the ``[E, d]`` gather, the padded
segment-sum, etc.  It is the checker for the CUDA path, never the product.

parity status: index construction pinned to the reference's own functions
(tests/golden/index_*.npz); propagation fwd/bwd pinned to the reference's own
model.py text executed over a numpy TF shim (tests/golden/modelref_*.npz) --
see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

PAD_ROWS = 100  # model.py:87  tf.pad(..., [[0,100],[0,0]])


# --------------------------------------------------------------------------
# index construction  (DataHandler.py:9-11, 47-69)
# --------------------------------------------------------------------------
def transpose(mat):
    """CSR of the transposed matrix.  Follows DataHandler.py:9-11
    (``csr_matrix(coo_matrix(mat).transpose())``): rows = items, column ids
    (= users) ascending inside a row, values carried along."""
    import scipy.sparse as sp

    return sp.csr_matrix(sp.coo_matrix(mat).transpose())


def trans_to_lsts(mat, norm=False):
    """Adjacency list of a scipy sparse matrix.  Follows DataHandler.py:47-69.

    Returns ``(indices int32 [E,2], data int32 [E], shape [R, C])`` with
      * indices row-major in COO order of the canonical CSR (DataHandler.py:49-50),
      * ``data`` = stored values cast to int32 (:51); with ``norm`` every value is
        multiplied (float64) by ``rowD[row] * colD[col]`` where
        ``rowD = 1/(sqrt(value_sum_row + 1e-8) + 1e-8)`` (:54-55) and written back
        into the *int32* array (:56-59), i.e. truncated toward zero,
      * an empty matrix replaced by the single fake edge ``[[0,0]]``, data ``[0]``
        (:66-68).
    The per-edge Python loop of the reference is vectorised; the arithmetic
    (operand order, dtypes, truncation) is the same.
    """
    import scipy.sparse as sp

    shape = [int(mat.shape[0]), int(mat.shape[1])]
    coo = sp.coo_matrix(mat)
    row = np.asarray(coo.row, dtype=np.int32)
    col = np.asarray(coo.col, dtype=np.int32)
    indices = np.stack([row, col], axis=1).astype(np.int32) if row.size else np.zeros((0, 2), np.int32)
    data = np.asarray(coo.data).astype(np.int32)
    if norm and data.size:
        row_sum = np.asarray(mat.sum(axis=1)).reshape(-1)   # int64 for intc matrices
        col_sum = np.asarray(mat.sum(axis=0)).reshape(-1)
        row_d = 1.0 / (np.sqrt(row_sum + 1e-8) + 1e-8)
        col_d = 1.0 / (np.sqrt(col_sum + 1e-8) + 1e-8)
        scaled = (data * row_d[row]) * col_d[col]           # float64, same operand order as :59
        data = scaled.astype(np.int32)                      # store into int32 array == C truncation
    if indices.shape[0] == 0:
        indices = np.array([[0, 0]], dtype=np.int32)
        data = np.array([0], dtype=np.int32)
    return indices, data, shape


def value_sum_degrees(mat):
    """int64 row / column sums of the stored values -- what DataHandler.py:54-55
    feeds into the (dead) normalisation."""
    r = np.asarray(mat.sum(axis=1)).reshape(-1).astype(np.int64)
    c = np.asarray(mat.sum(axis=0)).reshape(-1).astype(np.int64)
    return r, c


def structural_degrees(indices, n_rows, n_cols):
    """Edge counts per row / column of an adjacency list (what a real
    D^-1/2 A D^-1/2 needs)."""
    r = np.bincount(indices[:, 0], minlength=n_rows).astype(np.int32)
    c = np.bincount(indices[:, 1], minlength=n_cols).astype(np.int32)
    return r, c


def lightgcn_edge_weights(indices, n_rows, n_cols, dtype=np.float64):
    """Optional weighted mode (north_star's 1/sqrt(d_u d_i)); NOT what the
    reference executes.  Uses the reference's formula
    ``1/(sqrt(s+1e-8)+1e-8)`` (DataHandler.py:54-55) on structural degrees."""
    rdeg, cdeg = structural_degrees(indices, n_rows, n_cols)
    row_d = 1.0 / (np.sqrt(rdeg.astype(np.float64) + 1e-8) + 1e-8)
    col_d = 1.0 / (np.sqrt(cdeg.astype(np.float64) + 1e-8) + 1e-8)
    return (row_d[indices[:, 0]] * col_d[indices[:, 1]]).astype(dtype)


# --------------------------------------------------------------------------
# TF-1.14 op restatements
# --------------------------------------------------------------------------
def _segment_sum_sorted(data, ids):
    """tf.math.segment_sum: ids sorted non-decreasing, output ids[-1]+1 rows,
    skipped ids are zero rows.  Rejects unsorted ids like the TF CPU kernel."""
    ids = np.asarray(ids)
    if ids.size == 0:
        return np.zeros((0,) + data.shape[1:], dtype=data.dtype)
    if np.any(np.diff(ids) < 0):
        raise ValueError("segment ids are not increasing")  # TF: InvalidArgumentError
    n_seg = int(ids[-1]) + 1
    out = np.zeros((n_seg,) + data.shape[1:], dtype=data.dtype)
    starts = np.flatnonzero(np.concatenate(([True], ids[1:] != ids[:-1])))
    out[ids[starts]] = np.add.reduceat(data, starts, axis=0)
    return out


def _unsorted_segment_sum(data, ids, n_seg):
    """tf.math.unsorted_segment_sum (densification of the IndexedSlices
    gradient of GatherV2).  Stable sort keeps edge order inside a segment."""
    out = np.zeros((n_seg,) + data.shape[1:], dtype=data.dtype)
    if len(ids) == 0:
        return out
    order = np.argsort(ids, kind="stable")
    sid = np.asarray(ids)[order]
    starts = np.flatnonzero(np.concatenate(([True], sid[1:] != sid[:-1])))
    out[sid[starts]] = np.add.reduceat(data[order], starts, axis=0)
    return out


def leaky_relu(x, leaky):
    """Utils/NNLayers.py:135-136  ``tf.maximum(leaky*data, data)``."""
    return np.maximum(np.asarray(leaky, dtype=x.dtype) * x, x)


def leaky_relu_grad(z, g, leaky):
    """TF MaximumGrad with x=leaky*z, y=z: gradient goes to x where x >= y
    (SURVEY A.3).  For 0<leaky<1 that is z <= 0 (ties at z == 0 take the
    ``leaky`` branch)."""
    lz = np.asarray(leaky, dtype=z.dtype) * z
    to_x = lz >= z
    return np.where(to_x, np.asarray(leaky, dtype=g.dtype) * g, g)


# --------------------------------------------------------------------------
# messagePropagate  (model.py:80-92)
# --------------------------------------------------------------------------
def message_propagate_pre(srclats, indices, n_rows, strict_pad=False, edge_weight=None):
    """model.py:82-91 up to (not including) the activation: returns
    ``lat [n_rows, d]``.

    gather by col (:86) -> sorted segment_sum by row + pad 100 zero rows (:87)
    -> lookup range(n_rows) (:88-91).  TF-CPU raises when the padded tensor has
    fewer than ``n_rows`` rows (``strict_pad=True`` mirrors that); TF-GPU yields
    zeros, which is also what any run that works at all sees.
    ``edge_weight`` (None = reference-exact: values are ignored, model.py:84-86)
    multiplies each gathered row -- the optional LightGCN mode.
    """
    src = indices[:, 1]
    tgt = indices[:, 0]
    src_emb = srclats[src]                                   # GatherV2  [E, d]
    if edge_weight is not None:
        src_emb = src_emb * np.asarray(edge_weight, dtype=srclats.dtype)[:, None]
    seg = _segment_sum_sorted(src_emb, tgt)                  # [tgt[-1]+1, d]
    padded = np.concatenate([seg, np.zeros((PAD_ROWS, srclats.shape[1]), srclats.dtype)], axis=0)
    if padded.shape[0] < n_rows:
        if strict_pad:
            raise IndexError(
                f"indices[{padded.shape[0]}] = {padded.shape[0]} is not in [0, {padded.shape[0]})"
            )  # TF-CPU GatherV2 InvalidArgumentError
        padded = np.concatenate(
            [padded, np.zeros((n_rows - padded.shape[0], srclats.shape[1]), srclats.dtype)], axis=0
        )
    return padded[:n_rows]                                   # embedding_lookup(lat, range(R))


def message_propagate(srclats, indices, n_rows, leaky=0.5, strict_pad=False, edge_weight=None):
    """model.py:80-92 complete: ``Activate(lat, 'leakyRelu')`` (:92)."""
    return leaky_relu(message_propagate_pre(srclats, indices, n_rows, strict_pad, edge_weight), leaky)


# --------------------------------------------------------------------------
# the interval / layer loop  (model.py:118-134) and its TF-autodiff backward
# --------------------------------------------------------------------------
def propagate_forward(adj, tp_adj, u_embed, i_embed, n_layers, leaky=0.5, dtype=np.float64,
                      edge_weight=None, tp_edge_weight=None, strict_pad=False):
    """model.py:118-129 for all T intervals.

    adj[k] / tp_adj[k]: int32 [E_k, 2] adjacency lists of A_k (U x I) and A_k^T
    (as produced by trans_to_lsts, model.py:230-236).  u_embed [T,U,d], i_embed [T,I,d].
    Returns (user_vec [T,U,d], item_vec [T,I,d], tape) where tape keeps the
    pre-activations needed by ``propagate_backward``.
    """
    T = len(adj)
    U, I = u_embed.shape[1], i_embed.shape[1]
    u_embed = np.asarray(u_embed, dtype=dtype)
    i_embed = np.asarray(i_embed, dtype=dtype)
    user_vec = np.empty_like(u_embed)
    item_vec = np.empty_like(i_embed)
    tape = []
    for k in range(T):
        embs0 = [u_embed[k]]                                  # model.py:119
        embs1 = [i_embed[k]]                                  # model.py:120
        z0s, z1s = [], []
        ew = None if edge_weight is None else edge_weight[k]
        tew = None if tp_edge_weight is None else tp_edge_weight[k]
        for _ in range(n_layers):                             # model.py:121
            z0 = message_propagate_pre(embs1[-1], adj[k], U, strict_pad, ew)      # :122
            z1 = message_propagate_pre(embs0[-1], tp_adj[k], I, strict_pad, tew)  # :123
            embs0.append(leaky_relu(z0, leaky) + embs0[-1])   # :124
            embs1.append(leaky_relu(z1, leaky) + embs1[-1])   # :125
            z0s.append(z0)
            z1s.append(z1)
        acc0 = embs0[0]
        for e in embs0[1:]:                                   # tf.add_n, :126
            acc0 = acc0 + e
        acc1 = embs1[0]
        for e in embs1[1:]:                                   # :127
            acc1 = acc1 + e
        user_vec[k] = acc0
        item_vec[k] = acc1
        tape.append((z0s, z1s))
    return user_vec, item_vec, tape


def propagate_backward(adj, tp_adj, tape, g_user, g_item, n_layers, leaky=0.5, dtype=np.float64,
                       edge_weight=None, tp_edge_weight=None, pass_masks=None):
    """Reverse-mode sweep over the op list of ``propagate_forward`` exactly as
    TF1 autodiff would build it (implied by model.py:250): AddN fans the
    upstream out to every ``embs*[j]``; the residual add passes the gradient to
    both operands; MaximumGrad (tie rule in ``leaky_relu_grad``); the identity
    lookup / pad / segment_sum turn into a gather by target id; GatherV2's
    IndexedSlices gradient is densified by unsorted_segment_sum over source ids.

    g_user [T,U,d], g_item [T,I,d] dense upstream.  Returns
    (d_u_embed [T,U,d], d_i_embed [T,I,d]).

    pass_masks (optional): ``pass_masks[k][l] = (bool [U,d], bool [I,d])`` replaces the
    MaximumGrad decision (True = gradient passes unscaled).  LeakyReLU's derivative jumps at
    0, so an fp32 implementation whose |z| ~ 1e-7 pre-activation rounds to the other sign
    legitimately takes the other branch; parity tests therefore check the sign masks
    separately (mismatches only at near-ties) and the backward GIVEN the masks.
    """
    T = len(adj)
    g_user = np.asarray(g_user, dtype=dtype)
    g_item = np.asarray(g_item, dtype=dtype)
    U, I = g_user.shape[1], g_item.shape[1]
    d_u = np.empty_like(g_user)
    d_i = np.empty_like(g_item)
    for k in range(T):
        z0s, z1s = tape[k]
        ge0 = [g_user[k].copy() for _ in range(n_layers + 1)]   # AddN grad
        ge1 = [g_item[k].copy() for _ in range(n_layers + 1)]
        ew = None if edge_weight is None else np.asarray(edge_weight[k], dtype=dtype)
        tew = None if tp_edge_weight is None else np.asarray(tp_edge_weight[k], dtype=dtype)
        for l in range(n_layers - 1, -1, -1):
            ga0 = ge0[l + 1]                                    # grad of a_emb0 and of embs0[l]
            ga1 = ge1[l + 1]
            ge0[l] = ge0[l] + ga0
            ge1[l] = ge1[l] + ga1
            if pass_masks is None:
                dz0 = leaky_relu_grad(z0s[l], ga0, leaky)       # [U,d]
                dz1 = leaky_relu_grad(z1s[l], ga1, leaky)       # [I,d]
            else:
                lk = np.asarray(leaky, dtype=ga0.dtype)
                dz0 = np.where(pass_masks[k][l][0], ga0, lk * ga0)
                dz1 = np.where(pass_masks[k][l][1], ga1, lk * ga1)
            # user-side call: lat rows <- item table rows via adj[k]  (tgt=user, src=item)
            per_edge0 = dz0[adj[k][:, 0]]
            if ew is not None:
                per_edge0 = per_edge0 * ew[:, None]
            ge1[l] = ge1[l] + _unsorted_segment_sum(per_edge0, adj[k][:, 1], I)
            # item-side call: tgt=item, src=user via tp_adj[k]
            per_edge1 = dz1[tp_adj[k][:, 0]]
            if tew is not None:
                per_edge1 = per_edge1 * tew[:, None]
            ge0[l] = ge0[l] + _unsorted_segment_sum(per_edge1, tp_adj[k][:, 1], U)
        d_u[k] = ge0[0]
        d_i[k] = ge1[0]
    return d_u, d_i


def propagate(adj, tp_adj, u_embed, i_embed, g_user, g_item, n_layers, leaky=0.5, dtype=np.float64,
              edge_weight=None, tp_edge_weight=None):
    """Convenience: forward + backward, returns (user_vec, item_vec, d_u, d_i)."""
    uv, iv, tape = propagate_forward(adj, tp_adj, u_embed, i_embed, n_layers, leaky, dtype,
                                     edge_weight, tp_edge_weight)
    du, di = propagate_backward(adj, tp_adj, tape, g_user, g_item, n_layers, leaky, dtype,
                                edge_weight, tp_edge_weight)
    return uv, iv, du, di


# --------------------------------------------------------------------------
# sampled pair scores over the outputs  (model.py:171-173, 194-198)
# --------------------------------------------------------------------------
def pair_scores(u_rows, i_rows, uids, iids, leaky=0.5, activation=True):
    """``pckUlat = embedding_lookup(user_vector[k], suids[k])``, ``pckIlat = embedding_lookup(item_vector[k],
    siids[k])``, ``preds_one = reduce_sum(Activate(pckUlat * pckIlat, 'leakyRelu'), -1)`` (model.py:194-196);
    ``activation=False``: the plain ``reduce_sum(pckUlat * pckIlat, -1)`` of model.py:171-173.  Returns
    ``(scores [n], x [n, d])`` with ``x`` the products (the tape)."""
    x = u_rows[np.asarray(uids)] * i_rows[np.asarray(iids)]
    return (leaky_relu(x, leaky) if activation else x).sum(axis=-1), x


def pair_scores_backward(u_rows, i_rows, uids, iids, g_scores, leaky=0.5, activation=True):
    """TF autodiff of `pair_scores`: Sum -> broadcast, MaximumGrad (tie rule of `leaky_relu_grad`), Mul,
    then the two embedding_lookup gradients (IndexedSlices, summed per row like unsorted_segment_sum).
    Returns dense ``(d_u_rows, d_i_rows)``."""
    uids, iids = np.asarray(uids), np.asarray(iids)
    a, b = u_rows[uids], i_rows[iids]
    g = np.broadcast_to(np.asarray(g_scores, dtype=u_rows.dtype)[:, None], a.shape)
    gx = leaky_relu_grad(a * b, g, leaky) if activation else g
    return (_unsorted_segment_sum(gx * b, uids, u_rows.shape[0]), _unsorted_segment_sum(gx * a, iids, i_rows.shape[0]))


def pass_masks_from_tape(tape, leaky=0.5):
    """The MaximumGrad decisions of a forward tape: True where the gradient passes unscaled."""
    out = []
    for z0s, z1s in tape:
        out.append([(~(np.asarray(leaky, z0.dtype) * z0 >= z0), ~(np.asarray(leaky, z1.dtype) * z1 >= z1))
                    for z0, z1 in zip(z0s, z1s)])
    return out


def relerr(x, ref):
    """Parity metric of SURVEY 8(d): max|x-ref| / max|ref|."""
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    denom = float(np.max(np.abs(ref))) if ref.size else 0.0
    if denom == 0.0:
        return float(np.max(np.abs(x))) if x.size else 0.0
    return float(np.max(np.abs(x - ref)) / denom)
