"""numpy restatement of the reference's interval fusion (TEST INFRASTRUCTURE ONLY) -- the direct
CONSUMER of the propagation path's ``[R,T,d]`` output, restated so that the hand-off
layouts the path writes (``user_vector_tensor`` / the row-sharded receive slabs) have a checked reader.

Follows LIU-YUXI/SA-GNN ``model.py:135-155`` and ``Utils/attention.py:31-78``:

    BasicLSTMCell(latdim) over the T intervals (one cell object, shared by the user and the item call:
    model.py:136-146), DropoutWrapper on its outputs (identity at keepRate 1), tf.contrib.layers.layer_norm
    (model.py:152-153), MultiHeadSelfAttention.attention (attention.py:55-78) with
    ScaledDotProductAttention (attention.py:34-44: exp without max subtraction, + 1e-8 in the normaliser),
    tf.reduce_mean over the T axis (model.py:154-155).

PARITY UNPINNED: nothing of the reference executes here (TF 1.14 is not importable) and, unlike the
propagation, these lines were not run over the numpy shim; the per-op semantics below are the
builder's reading of TF 1.14 (BasicLSTMCell: gate order i, j, f, o, forget_bias 1.0; layer_norm:
begin_norm_axis=1, begin_params_axis=-1, epsilon 1e-12; tf.layers.dense: kernel + bias).
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
