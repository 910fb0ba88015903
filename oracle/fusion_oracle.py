"""numpy restatement of the reference's interval fusion (TEST INFRASTRUCTURE ONLY) -- the direct
CONSUMER of the propagation path's ``[R,T,d]`` output, restated so that the hand-off
layouts the path writes (``user_vector_tensor`` / the row-sharded receive slabs) have a checked reader.

Follows LIU-YUXI/SA-GNN ``model.py:135-155`` and ``Utils/attention.py:31-78``:

    BasicLSTMCell(latdim) over the T intervals (one cell object, shared by the user and the item call:
    model.py:136-146), DropoutWrapper on its outputs (identity at keepRate 1), tf.contrib.layers.layer_norm
    (model.py:152-153), MultiHeadSelfAttention.attention (attention.py:55-78) with
    ScaledDotProductAttention (attention.py:34-44: exp without max subtraction, + 1e-8 in the normaliser),
    tf.reduce_mean over the T axis (model.py:154-155).

PINNED TO THE REFERENCE TEXT (round 2): ``tests/golden/make_golden_downstream.py`` executes model.py:111-112 and
133-203 as they stand (with the reference's own Utils/attention.py and Utils/NNLayers.py) over the
numpy TF stand-in and commits ``tests/golden/downstream_*.npz``; ``tests/test_fusion.py`` compares every function
below with those fixtures.  Wiring, the shared LSTM cell, per-side Q / K / V kernels, exp-normalised attention,
axes, stop_gradient placement and the positive | negative slicing therefore come from the reference; what stays
the builder's statement is the per-op semantics of the TF 1.14 kernels inside the stand-in (BasicLSTMCell: gate
order i, j, f, o, forget_bias 1.0; layer_norm: begin_norm_axis=1, begin_params_axis=-1, epsilon 1e-12;
tf.layers.dense: kernel + bias).
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def basic_lstm(x, kernel, bias, forget_bias=1.0):
    """tf.nn.dynamic_rnn(BasicLSTMCell(d), x [R,T,d], zero state) -> outputs [R,T,d]   (model.py:136-146).
    kernel [2d, 4d] multiplies concat([x_t, h], 1); gates split as i, j, f, o."""
    R, T, d = x.shape
    h = np.zeros((R, d), x.dtype)
    c = np.zeros((R, d), x.dtype)
    out = np.empty_like(x)
    for t in range(T):
        g = np.concatenate([x[:, t], h], axis=1) @ kernel + bias
        i, j, f, o = np.split(g, 4, axis=1)
        c = c * _sigmoid(f + forget_bias) + _sigmoid(i) * np.tanh(j)
        h = np.tanh(c) * _sigmoid(o)
        out[:, t] = h
    return out


def layer_norm(x, gamma, beta, eps=1e-12):
    """tf.contrib.layers.layer_norm(x [R,T,d]): moments over axes (1, 2), scale / shift over the last axis
    (model.py:152-153)."""
    mean = x.mean(axis=(1, 2), keepdims=True)
    var = x.var(axis=(1, 2), keepdims=True)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta


def multihead_self_attention(x, wq, bq, wk, bk, wv, bv, heads):
    """MultiHeadSelfAttention(d, heads).attention(x [R,T,d])   (Utils/attention.py:55-78, 34-44)."""
    R, T, d = x.shape
    dk = d // heads
    split = lambda y: y.reshape(R, T, heads, dk).transpose(0, 2, 1, 3)         # [R, heads, T, dk]
    q, k, v = split(x @ wq + bq), split(x @ wk + bk), split(x @ wv + bv)
    scores = np.exp(q @ k.transpose(0, 1, 3, 2) / np.sqrt(dk))                  # no max subtraction
    attn = scores / (scores.sum(axis=-1, keepdims=True) + 1e-8)
    ctx = attn @ v                                                              # [R, heads, T, dk]
    return ctx.transpose(0, 2, 1, 3).reshape(R, T, heads * dk)


def interval_fusion(x, p, heads):
    """model.py:135-155 for one side: x [R,T,d] -> final vector [R,d].  p: dict of that side's parameters
    (lstm_kernel / lstm_bias shared by both sides; ln_gamma, ln_beta, wq, bq, wk, bk, wv, bv per side)."""
    h = basic_lstm(x, p["lstm_kernel"], p["lstm_bias"])
    n = layer_norm(h, p["ln_gamma"], p["ln_beta"])
    a = multihead_self_attention(n, p["wq"], p["bq"], p["wk"], p["bk"], p["wv"], p["bv"], heads)
    return a.mean(axis=1)


def _lrelu(x, leaky):
    return np.maximum(leaky * x, x)


def sequence_attention(final_item, sequence, mask, pos_embed, ln_seq, ln_pos, layers, heads, leaky):
    """model.py:111-112,157-168: the user's item sequence (right-aligned ids ``sequence`` [B,P], 0 / 1 ``mask`` [B,P])
    collapses to ONE vector per user before any attention: ``mask @ final_item[sequence]`` ([B,1,P] x [B,P,d]),
    layer-normed, plus the layer-normed masked sum of the position embeddings; then ``len(layers)`` rounds of
    ``x = lrelu(MHSA(layer_norm(x))) + x`` over that length-1 axis, and ``att_user = sum over it`` -> [B,d].
    ``ln_seq`` / ``ln_pos`` = (gamma, beta); ``layers`` = list of dicts ln_gamma, ln_beta, wq, bq, wk, bk, wv, bv."""
    m = mask[:, None, :]
    x = layer_norm(m @ final_item[sequence], *ln_seq) + layer_norm(m @ pos_embed[None, :, :], *ln_pos)
    for p in layers:
        a = multihead_self_attention(layer_norm(x, p["ln_gamma"], p["ln_beta"]), p["wq"], p["bq"], p["wk"], p["bk"],
                                     p["wv"], p["bv"], heads)
        x = _lrelu(a, leaky) + x
    return x.sum(axis=1)


def predictions(final_user, final_item, att_user, uids, iids, u_locs_seq, leaky):
    """model.py:169-173: ``preds = <final_user[uids], final_item[iids]> + sum(lrelu(att_user[uLocs_seq]) *
    final_item[iids])`` (``iEmbed_att`` IS ``final_item_vector``, model.py:156)."""
    it = final_item[iids]
    return (final_user[uids] * it).sum(-1) + (_lrelu(att_user[u_locs_seq], leaky) * it).sum(-1)


def meta_user_weight(final_user, user_vector, w2, b2, w3, b3, leaky):
    """model.py:178-184: per interval k, ``meta1 = concat([final * uv_k, final, uv_k], -1)``, ``meta2 =
    lrelu(meta1 @ W2 + b2)``, ``weight_k = sigmoid(meta2 @ W3 + b3)`` squeezed to [U]; stacked to [T,U].
    The same two FC layers (``reuse=True``) serve every interval."""
    out = []
    for k in range(user_vector.shape[0]):
        m1 = np.concatenate([final_user * user_vector[k], final_user, user_vector[k]], axis=-1)
        m2 = _lrelu(m1 @ w2 + b2, leaky)
        out.append(_sigmoid(m2 @ w3 + b3)[:, 0])
    return np.stack(out, axis=0)


def ssl_hinge(final_user, final_item, user_vector, item_vector, user_weight, suids, siids, leaky):
    """model.py:185-203: for every interval, the first half of (suids, siids) are the positive pairs, the second
    half the negatives; ``S = w[pos users] * s_final[pos] - w[neg users] * s_final[neg]`` with the FINAL-vector
    scores under stop_gradient, ``preds_one`` = the interval's own pair scores, loss = sum max(0, 1 - S * (pos - neg)).
    Returns (sslloss, [preds_one per interval])."""
    loss, preds = 0.0, []
    for k in range(user_vector.shape[0]):
        su, si = np.asarray(suids[k]), np.asarray(siids[k])
        n = su.shape[0] // 2
        s_final = _lrelu(final_user[su] * final_item[si], leaky).sum(axis=-1)
        w = user_weight[k][su]
        S = w[:n] * s_final[:n] - w[n:] * s_final[n:]
        p1 = _lrelu(user_vector[k][su] * item_vector[k][si], leaky).sum(axis=-1)
        loss = loss + np.maximum(0.0, 1.0 - S * (p1[:n] - p1[n:])).sum()
        preds.append(p1)
    return loss, preds
