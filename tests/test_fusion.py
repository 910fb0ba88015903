"""Interval fusion (the propagation path's consumer): torch module vs the numpy restatement
of model.py:135-155 / Utils/attention.py:31-78, and the slab -> [R,T,d] assembly of a row-sharded consumer."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from sagnn_b200.fusion import IntervalFusion, slabs_to_rtd


def _module(d, heads, seed):
    m = IntervalFusion(d, heads=heads, dtype=torch.float64, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():                      # biases / layer-norm parameters off their trivial initial values
            p.add_(torch.randn(p.shape, generator=g, dtype=torch.float64) * 0.1)
    return m


@pytest.mark.parametrize("d,heads,T,R", [(64, 16, 3, 23), (32, 8, 5, 40), (128, 16, 1, 9)])
def test_fusion_module_matches_oracle(d, heads, T, R):
    m = _module(d, heads, seed=d + T)
    x = torch.randn((R, T, d), dtype=torch.float64, generator=torch.Generator().manual_seed(7))
    for side in ("user", "item"):
        got = m.fuse(x, side).detach().numpy()
        ref = fo.interval_fusion(x.numpy(), {k: v.detach().numpy() for k, v in m.side_params(side).items()}, heads)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    u, i = m(x, x[: R // 2])
    assert u.shape == (R, d) and i.shape == (R // 2, d)


def test_lstm_single_step_closed_form():
    """T = 1 from a zero state: c = sigmoid(i) * tanh(j), h = tanh(c) * sigmoid(o); the forget gate (and its
    bias of 1.0) cannot matter (BasicLSTMCell, gate order i, j, f, o)."""
    d = 4
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 1, d))
    k = rng.standard_normal((2 * d, 4 * d)); b = rng.standard_normal(4 * d)
    g = x[:, 0] @ k[:d] + b
    i, j, f, o = np.split(g, 4, axis=1)
    sig = lambda z: 1 / (1 + np.exp(-z))
    want = np.tanh(sig(i) * np.tanh(j)) * sig(o)
    np.testing.assert_allclose(fo.basic_lstm(x, k, b)[:, 0], want, rtol=1e-13)
    np.testing.assert_allclose(fo.basic_lstm(x, k, b, forget_bias=5.0)[:, 0], want, rtol=1e-13)


def test_layer_norm_is_over_intervals_and_features():
    x = np.random.default_rng(1).standard_normal((5, 3, 8))
    y = fo.layer_norm(x, np.ones(8), np.zeros(8))
    np.testing.assert_allclose(y.mean(axis=(1, 2)), 0, atol=1e-12)
    np.testing.assert_allclose(y.var(axis=(1, 2)), 1, rtol=1e-9)


def test_attention_rows_are_normalised_by_sum_plus_eps():
    x = np.random.default_rng(2).standard_normal((2, 4, 8))
    eye, z = np.eye(8), np.zeros(8)
    out = fo.multihead_self_attention(x, eye, z, eye, z, eye, z, heads=2)
    q = x.reshape(2, 4, 2, 4).transpose(0, 2, 1, 3)
    s = np.exp(q @ q.transpose(0, 1, 3, 2) / 2.0)
    want = ((s / (s.sum(-1, keepdims=True) + 1e-8)) @ q).transpose(0, 2, 1, 3).reshape(2, 4, 8)
    np.testing.assert_allclose(out, want, rtol=1e-13)


def test_slabs_to_rtd_orders_intervals_by_owner():
    """Receive slabs are [source rank, blk, T_local, d] in every source rank's LOCAL interval order; the LSTM
    needs true interval order."""
    owners = [1, 0, 1, 0, 0]                     # rank 0 owns intervals 1, 3, 4; rank 1 owns 0, 2
    blk, d = 6, 4
    full = torch.arange(5 * blk * d, dtype=torch.float32).reshape(5, blk, d)     # [T, blk, d]
    slabs = torch.zeros((2, blk, 3, d))
    for src in range(2):
        for j, k in enumerate([k for k in range(5) if owners[k] == src]):
            slabs[src, :, j] = full[k]
    x = slabs_to_rtd(slabs, owners, rank_rows=5)
    assert x.shape == (5, 5, d) and torch.equal(x, full.transpose(0, 1)[:5])
