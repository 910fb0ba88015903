"""Interval fusion: the CONSUMER of the propagation path (LIU-YUXI/SA-GNN
``model.py:135-155``, ``Utils/attention.py:31-78``), so that the hand-off layouts the path writes have a reader:

  * ``IntervalFusion``: BasicLSTMCell over the T intervals -> layer norm -> multi-head self-attention ->
    mean over T, on the ``[R,T,d]`` tensors ``propagate(..., layout="rtd")`` returns
    (``user_vector_tensor`` / ``item_vector_tensor``, model.py:133-134).  Dense framework-side work (GEMMs and
    pointwise ops through torch / cuBLAS), NOT part of the sparse hot path and not hand-written kernels;
    its autograd produces the dense ``[R,T,d]`` upstream that ``sagnn_propagate_bwd_ex`` consumes.
  * ``slabs_to_rtd``: assembles a row-sharded consumer's input from the receive slabs
    ``[source rank, row block, T_local, d]`` that ``sagnn_propagate_fwd_scatter`` (peer stores from the epilogue)
    or the copy-engine hand-off fill: rank j holds row block j of EVERY interval, which is all an LSTM over T needs.

Same semantics as TF 1.14 for the ops involved (gate order i, j, f, o with forget_bias 1.0; layer norm over
(T, d) with epsilon 1e-12; exp-normalised attention with 1e-8 in the denominator); checked against
``oracle/fusion_oracle.py`` and, through ``tests/golden/downstream_*.npz``, against the reference's own
model.py:133-156 / Utils/attention.py text executed over a numpy TF stand-in (``tests/test_fusion.py``).

  * ``SequenceAttention``: the sequence branch of the prediction (model.py:111-112, 157-168, 173).
  * ``SslHead``: the meta-weight network and the weighted hinge of the SSL objective (model.py:174-203) on top of
    the pair scores ``sagnn_b200.pair_scores`` gathers from the path's outputs (dense framework-side ops as well).
"""
from __future__ import annotations

import math

import torch


def _layer_norm(x, gamma, beta):
    """tf.contrib.layers.layer_norm on [R,T,d]: moments over (T, d), scale / shift over d, epsilon 1e-12."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-12) * gamma + beta


def _mhsa(n, p, heads):
    """MultiHeadSelfAttention.attention (Utils/attention.py:46-78, 31-44): n [R,T,d] -> [R,T,d]; p: wq, bq, wk, bk, wv, bv."""
    R, T, d = n.shape
    dk = d // heads
    split = lambda y: y.reshape(R, T, heads, dk).permute(0, 2, 1, 3)
    q, k, v = split(n @ p["wq"] + p["bq"]), split(n @ p["wk"] + p["bk"]), split(n @ p["wv"] + p["bv"])
    scores = torch.exp(q @ k.transpose(-1, -2) / math.sqrt(dk))
    attn = scores / (scores.sum(dim=-1, keepdim=True) + 1e-8)
    return (attn @ v).permute(0, 2, 1, 3).reshape(R, T, d)


class IntervalFusion(torch.nn.Module):
    def __init__(self, d, heads=16, dtype=torch.float32, device=None, seed=0):
        super().__init__()
        if d % heads:
            raise ValueError("latdim must be a multiple of the number of attention heads (attention.py:49)")
        self.d, self.heads = d, heads
        g = torch.Generator().manual_seed(seed)

        def xavier(rows, cols):
            a = math.sqrt(6.0 / (rows + cols))
            return torch.nn.Parameter(((torch.rand((rows, cols), generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype=dtype, device=device))

        zeros = lambda n: torch.nn.Parameter(torch.zeros(n, dtype=dtype, device=device))
        self.lstm_kernel, self.lstm_bias = xavier(2 * d, 4 * d), zeros(4 * d)    # one cell for both sides (model.py:141-146)
        for side in ("user", "item"):
            setattr(self, side + "_ln_gamma", torch.nn.Parameter(torch.ones(d, dtype=dtype, device=device)))
            setattr(self, side + "_ln_beta", zeros(d))
            for n in ("q", "k", "v"):
                setattr(self, "%s_w%s" % (side, n), xavier(d, d))
                setattr(self, "%s_b%s" % (side, n), zeros(d))

    def side_params(self, side):
        names = {"lstm_kernel": "lstm_kernel", "lstm_bias": "lstm_bias", "ln_gamma": side + "_ln_gamma",
                 "ln_beta": side + "_ln_beta"}
        names.update({k + n: "%s_%s%s" % (side, k, n) for n in ("q", "k", "v") for k in ("w", "b")})
        return {k: getattr(self, v) for k, v in names.items()}

    def lstm(self, x):
        R, T, d = x.shape
        h = x.new_zeros((R, d))
        c = x.new_zeros((R, d))
        outs = []
        for t in range(T):
            g = torch.cat([x[:, t], h], dim=1) @ self.lstm_kernel + self.lstm_bias
            i, j, f, o = g.split(d, dim=1)
            c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
            h = torch.tanh(c) * torch.sigmoid(o)
            outs.append(h)
        return torch.stack(outs, dim=1)

    def fuse(self, x, side):
        """x [R,T,d] -> [R,d]   (model.py:146-155 for one side)."""
        p = self.side_params(side)
        R, T, d = x.shape
        n = _layer_norm(self.lstm(x), p["ln_gamma"], p["ln_beta"])
        return _mhsa(n, p, self.heads).mean(dim=1)

    def forward(self, user_rtd, item_rtd):
        return self.fuse(user_rtd, "user"), self.fuse(item_rtd, "item")


class SequenceAttention(torch.nn.Module):
    """The sequence branch of the prediction (model.py:111-112, 157-168, 173): a user's item sequence is collapsed to
    the masked SUM of the final item vectors (plus the masked sum of position embeddings, both layer-normed), refined
    by ``att_layers`` rounds of ``x = lrelu(MHSA(layer_norm(x))) + x`` over that length-1 axis -> ``att_user`` [B,d];
    ``predict`` adds its term to the dot product of the final vectors (model.py:169-173)."""

    def __init__(self, d, heads=16, att_layers=4, pos_length=200, leaky=0.5, dtype=torch.float32, device=None, seed=0):
        super().__init__()
        if d % heads:
            raise ValueError("latdim must be a multiple of the number of attention heads (attention.py:49)")
        self.d, self.heads, self.att_layers, self.leaky = d, heads, att_layers, leaky
        g = torch.Generator().manual_seed(seed)
        P = lambda t: torch.nn.Parameter(t.to(dtype=dtype, device=device))

        def xavier(rows, cols):
            a = math.sqrt(6.0 / (rows + cols))
            return P((torch.rand((rows, cols), generator=g, dtype=torch.float64) * 2 - 1) * a)

        self.pos_embed = xavier(pos_length, d)
        for name in ["seq_ln", "pos_ln"] + ["l%d_ln" % l for l in range(att_layers)]:
            setattr(self, name + "_gamma", P(torch.ones(d)))
            setattr(self, name + "_beta", P(torch.zeros(d)))
        for l in range(att_layers):
            for n in ("q", "k", "v"):
                setattr(self, "l%d_w%s" % (l, n), xavier(d, d))
                setattr(self, "l%d_b%s" % (l, n), P(torch.zeros(d)))

    def layer_params(self, l):
        out = {"ln_gamma": getattr(self, "l%d_ln_gamma" % l), "ln_beta": getattr(self, "l%d_ln_beta" % l)}
        out.update({k + n: getattr(self, "l%d_%s%s" % (l, k, n)) for n in ("q", "k", "v") for k in ("w", "b")})
        return out

    def forward(self, final_item, sequence, mask):
        """final_item [I,d], sequence [B,P] item ids (right-aligned, 0-padded), mask [B,P] 0 / 1 -> att_user [B,d]."""
        m = mask.to(final_item.dtype).unsqueeze(1)                                    # [B,1,P]
        x = _layer_norm(m @ final_item[sequence], self.seq_ln_gamma, self.seq_ln_beta) + \
            _layer_norm(m @ self.pos_embed.unsqueeze(0), self.pos_ln_gamma, self.pos_ln_beta)
        for l in range(self.att_layers):
            p = self.layer_params(l)
            a = _mhsa(_layer_norm(x, p["ln_gamma"], p["ln_beta"]), p, self.heads)
            x = torch.maximum(self.leaky * a, a) + x
        return x.sum(dim=1)

    def predict(self, final_user, final_item, att_user, uids, iids, u_locs_seq):
        """model.py:169-173.  (On the GPU the first term is ``sagnn_b200.pair_scores(..., activation=None)``.)"""
        it = final_item[iids]
        s = att_user[u_locs_seq]
        return (final_user[uids] * it).sum(-1) + (torch.maximum(self.leaky * s, s) * it).sum(-1)


def prediction_hinge(preds):
    """model.py:241-244: ``preds`` holds the positives of all users, then their negatives (the sampler's order);
    ``preLoss = mean(max(0, 1 - (pos - neg)))``."""
    n = preds.shape[0] // 2
    return torch.clamp_min(1.0 - (preds[:n] - preds[n:]), 0.0).mean()


class SslHead(torch.nn.Module):
    """model.py:174-203 around the gathered pair scores: ``user_weight`` (two FC layers shared by all intervals:
    ``meta2`` [3d, ssldim] + bias with LeakyReLU, ``meta3`` [ssldim, 1] + bias with sigmoid) and the hinge
    ``sum max(0, 1 - S * (pos - neg))`` with ``S = w[pos] * s_final[pos] - w[neg] * s_final[neg]`` (final-vector
    scores detached, like the reference's ``tf.stop_gradient``)."""

    def __init__(self, d, ssldim=32, leaky=0.5, dtype=torch.float32, device=None, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)

        def xavier(rows, cols):
            a = math.sqrt(6.0 / (rows + cols))
            return torch.nn.Parameter(((torch.rand((rows, cols), generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype=dtype, device=device))

        self.leaky = leaky
        self.meta2, self.meta2_bias = xavier(3 * d, ssldim), torch.nn.Parameter(torch.zeros(ssldim, dtype=dtype, device=device))
        self.meta3, self.meta3_bias = xavier(ssldim, 1), torch.nn.Parameter(torch.zeros(1, dtype=dtype, device=device))

    def user_weight(self, final_user, user_vector):
        """final_user [U,d], user_vector [T,U,d] (the path's output) -> [T,U]   (model.py:178-184)."""
        f = final_user.unsqueeze(0).expand_as(user_vector)
        m1 = torch.cat([f * user_vector, f, user_vector], dim=-1)
        z = m1 @ self.meta2 + self.meta2_bias
        m2 = torch.maximum(self.leaky * z, z)                     # Activate(..., 'leakyRelu'), Utils/NNLayers.py:135-136
        return torch.sigmoid(m2 @ self.meta3 + self.meta3_bias).squeeze(-1)

    @staticmethod
    def hinge(weight_at_suids, final_scores, interval_scores):
        """One interval (model.py:186-202): all three are [2n] in the sampler's positives | negatives order;
        ``final_scores`` = pair scores on the final vectors (detached here), ``interval_scores`` = ``preds_one``."""
        n = interval_scores.shape[0] // 2
        sf = final_scores.detach()
        S = weight_at_suids[:n] * sf[:n] - weight_at_suids[n:] * sf[n:]
        return torch.clamp_min(1.0 - S * (interval_scores[:n] - interval_scores[n:]), 0.0).sum()


def slabs_to_rtd(slabs, owners, rank_rows=None):
    """Receive slabs ``[world, blk, T_local, d]`` (slab r = my row block of source rank r's intervals, in that
    rank's local interval order) -> ``[blk, T, d]`` in true interval order.  ``owners[k]`` = rank that owns
    interval k (``dist.assign_intervals``).  ``rank_rows``: keep only the first rows of the (padded) block."""
    world = slabs.shape[0]
    local = [[k for k, r in enumerate(owners) if r == src] for src in range(world)]
    T = len(owners)
    cols = [None] * T
    for src in range(world):
        for j, k in enumerate(local[src]):
            cols[k] = slabs[src, :, j]
    x = torch.stack(cols, dim=1)
    return x if rank_rows is None else x[:rank_rows]
