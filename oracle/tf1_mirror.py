"""torch-CPU op-by-op mirror of the TF1 graph of the path (TEST INFRASTRUCTURE ONLY).

This is the timed *CPU baseline* ("TF1-graph restatement, not TF": BASELINE.md
section 2, b1/b2) and an independent cross-check of ``propagate_oracle``: the
forward is written with the same op sequence ``model.py:80-92,118-129``
executes (GatherV2 -> SegmentSum -> Pad -> GatherV2(range) -> Mul/Maximum ->
Add -> AddN) with every intermediate materialised, and the backward is left to
torch autograd the way the reference leaves it to TF autodiff
(``model.py:250``).  The only hand-written gradient is the LeakyReLU tie rule
(TF MaximumGrad sends ties to the ``leaky*x`` operand, torch.maximum splits
them 50/50).

Never imported by the product path.  parity status: UNPINNED (see oracle/__init__.py).
"""
from __future__ import annotations

import torch

PAD_ROWS = 100  # model.py:87


class _TFLeakyRelu(torch.autograd.Function):
    """tf.maximum(leaky*x, x)  (Utils/NNLayers.py:135-136) with TF's MaximumGrad."""

    @staticmethod
    def forward(ctx, x, leaky):
        lx = x * leaky                       # Mul
        ctx.save_for_backward(lx >= x)       # MaximumGrad mask: to the first operand where a >= b
        ctx.leaky = leaky
        return torch.maximum(lx, x)          # Maximum

    @staticmethod
    def backward(ctx, g):
        (to_x,) = ctx.saved_tensors
        return torch.where(to_x, g * ctx.leaky, g), None


def message_propagate(srclats, indices, n_rows, leaky, edge_weight=None):
    """model.py:80-92.  indices: int64 [E,2] (SparseTensor.indices are int64 in-graph)."""
    src = indices[:, 1]
    tgt = indices[:, 0]
    src_emb = srclats.index_select(0, src)                       # GatherV2, materialises [E,d]
    if edge_weight is not None:
        src_emb = src_emb * edge_weight[:, None]
    n_seg = int(tgt[-1]) + 1
    seg = torch.zeros((n_seg, srclats.shape[1]), dtype=srclats.dtype)
    seg = seg.index_add(0, tgt, src_emb)                         # SegmentSum (sorted ids)
    lat = torch.nn.functional.pad(seg, (0, 0, 0, PAD_ROWS))      # Pad 100 zero rows
    if lat.shape[0] < n_rows:                                    # TF-GPU behaviour: zeros
        lat = torch.nn.functional.pad(lat, (0, 0, 0, n_rows - lat.shape[0]))
    lat = lat.index_select(0, torch.arange(n_rows))              # GatherV2 with range(R)
    return _TFLeakyRelu.apply(lat, leaky)


def propagate_forward(adj, tp_adj, u_embed, i_embed, n_layers, leaky=0.5,
                      edge_weight=None, tp_edge_weight=None):
    """model.py:118-134.  adj/tp_adj: lists of int64 [E_k,2] tensors."""
    users, items = [], []
    for k in range(len(adj)):
        embs0 = [u_embed[k]]
        embs1 = [i_embed[k]]
        ew = None if edge_weight is None else edge_weight[k]
        tew = None if tp_edge_weight is None else tp_edge_weight[k]
        for _ in range(n_layers):
            a0 = message_propagate(embs1[-1], adj[k], u_embed.shape[1], leaky, ew)
            a1 = message_propagate(embs0[-1], tp_adj[k], i_embed.shape[1], leaky, tew)
            embs0.append(a0 + embs0[-1])
            embs1.append(a1 + embs1[-1])
        users.append(torch.stack(embs0).sum(0))                  # AddN
        items.append(torch.stack(embs1).sum(0))
    return torch.stack(users, 0), torch.stack(items, 0)         # Pack


def propagate(adj, tp_adj, u_embed, i_embed, g_user, g_item, n_layers, leaky=0.5,
              edge_weight=None, tp_edge_weight=None):
    """fwd + autograd bwd with dense upstream; returns (user_vec, item_vec, dU, dI)."""
    u = u_embed.detach().clone().requires_grad_(True)
    i = i_embed.detach().clone().requires_grad_(True)
    uv, iv = propagate_forward(adj, tp_adj, u, i, n_layers, leaky, edge_weight, tp_edge_weight)
    torch.autograd.backward([uv, iv], [g_user, g_item])
    return uv.detach(), iv.detach(), u.grad, i.grad
