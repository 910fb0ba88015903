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


# ---- pinned to the reference's own text: model.py:111-112, 133-203 executed over the numpy TF stand-in
#      (tests/golden/make_golden_downstream.py -> downstream_*.npz) ------------------------------------------------
import glob
import os

_DOWNSTREAM = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "downstream_*.npz")))


def _layout(fx):
    """Variables in the order the reference text creates them (checked against the recorded names): posEmbed; the ONE
    LSTM cell (kernel, bias); per side LayerNorm (beta, gamma) + Q, K, V dense (kernel, bias each), user first; the
    sequence branch: LayerNorm of the item sum, LayerNorm of the position sum, per attention layer LayerNorm + Q, K, V;
    meta2, meta2Bias, meta3, meta3Bias."""
    names = [str(n).split("#")[0] for n in fx["var_names"]]
    L = int(fx["att_layer"])
    ln, qkv = ["LayerNorm/beta", "LayerNorm/gamma"], ["dense/kernel", "dense/bias"] * 3
    want = ["posEmbed", "basic_lstm_cell/kernel", "basic_lstm_cell/bias"] + (ln + qkv) * 2 + ln + ln + (ln + qkv) * L + \
        ["meta2", "meta2Bias", "meta3", "meta3Bias"]
    assert names == want
    return dict(pos=0, lstm=1, user=3, item=11, seq_ln=19, pos_ln=21, layers=[23 + 8 * l for l in range(L)], meta=23 + 8 * L)


def _block(fx, o, dtype=np.float64):
    """LayerNorm + Q, K, V starting at variable o."""
    v = lambda j: fx["var%02d" % j].astype(dtype)
    return dict(ln_beta=v(o), ln_gamma=v(o + 1), wq=v(o + 2), bq=v(o + 3), wk=v(o + 4), bk=v(o + 5), wv=v(o + 6), bv=v(o + 7))


def _side_params(fx, side, dtype=np.float64):
    lay = _layout(fx)
    v = lambda j: fx["var%02d" % j].astype(dtype)
    return dict(lstm_kernel=v(lay["lstm"]), lstm_bias=v(lay["lstm"] + 1), **_block(fx, lay[side], dtype))


def _meta(fx, dtype=np.float64):
    o = _layout(fx)["meta"]
    return [fx["var%02d" % (o + j)].astype(dtype) for j in range(4)]


def test_downstream_fixtures_present():
    assert len(_DOWNSTREAM) >= 2


@pytest.mark.parametrize("path", _DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in _DOWNSTREAM])
def test_fusion_oracle_matches_reference_model_py_executed(path):
    fx = np.load(path)
    heads = int(fx["heads"])
    for side, x, want in (("user", fx["user_vector"], fx["final_user_vector"]), ("item", fx["item_vector"], fx["final_item_vector"])):
        got = fo.interval_fusion(x.astype(np.float64).transpose(1, 0, 2), _side_params(fx, side), heads)
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("path", _DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in _DOWNSTREAM])
def test_fusion_module_matches_reference_model_py_executed(path):
    """The product's consumer module (forward in fp64 and fp32, autograd input gradient) against the executed text
    and its fp64 finite differences."""
    fx = np.load(path)
    d, heads = int(fx["d"]), int(fx["heads"])
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 2e-5)):
        m = IntervalFusion(d, heads=heads, dtype=dtype)
        with torch.no_grad():
            for side in ("user", "item"):
                for k, val in _side_params(fx, side).items():
                    m.side_params(side)[k].copy_(torch.from_numpy(val))
        u = torch.from_numpy(fx["user_vector"]).to(dtype).transpose(0, 1).contiguous().requires_grad_(True)
        i = torch.from_numpy(fx["item_vector"]).to(dtype).transpose(0, 1).contiguous().requires_grad_(True)
        fu, fi = m(u, i)
        scale = np.abs(fx["final_user_vector"]).max()
        assert np.abs(fu.detach().numpy() - fx["final_user_vector"]).max() <= tol * scale
        assert np.abs(fi.detach().numpy() - fx["final_item_vector"]).max() <= tol * scale
        if dtype == torch.float64:
            ((fu * torch.from_numpy(fx["w_user"])).sum() + (fi * torch.from_numpy(fx["w_item"])).sum()).backward()
            for got, want in ((u.grad, fx["d_user_vector"]), (i.grad, fx["d_item_vector"])):
                got = got.transpose(0, 1).numpy()           # fixtures are [T,R,d]
                assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


@pytest.mark.parametrize("path", _DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in _DOWNSTREAM])
def test_pair_score_and_ssl_oracles_match_reference_model_py_executed(path):
    """N2: ``preds`` (model.py:169-172), ``preds_one`` / ``user_weight`` / ``sslloss`` (model.py:174-203)."""
    from oracle import propagate_oracle as po
    fx = np.load(path)
    T, leaky = int(fx["T"]), float(fx["leaky"])
    uv, iv = fx["user_vector"].astype(np.float64), fx["item_vector"].astype(np.float64)
    fu, fi = fx["final_user_vector"], fx["final_item_vector"]
    s, _ = po.pair_scores(fu, fi, fx["uids"], fx["iids"], leaky, activation=False)
    np.testing.assert_allclose(s, fx["preds_dot"], rtol=1e-12, atol=1e-12)
    for k in range(T):
        s, _ = po.pair_scores(uv[k], iv[k], fx["suids%d" % k], fx["siids%d" % k], leaky, activation=True)
        np.testing.assert_allclose(s, fx["preds_one%d" % k], rtol=1e-12, atol=1e-12)
    w = fo.meta_user_weight(fu, uv, *_meta(fx), leaky)
    np.testing.assert_allclose(w, fx["user_weight"], rtol=1e-12, atol=1e-14)
    loss, p1 = fo.ssl_hinge(fu, fi, uv, iv, w, [fx["suids%d" % k] for k in range(T)], [fx["siids%d" % k] for k in range(T)], leaky)
    np.testing.assert_allclose(loss, float(fx["sslloss"]), rtol=1e-12)


@pytest.mark.parametrize("path", _DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in _DOWNSTREAM])
def test_ssl_head_matches_reference_model_py_executed(path):
    """The product's SslHead (meta-weight network + hinge) fed with gathered scores, against the executed text."""
    from sagnn_b200.fusion import SslHead
    fx = np.load(path)
    T, d, leaky = int(fx["T"]), int(fx["d"]), float(fx["leaky"])
    head = SslHead(d, ssldim=int(fx["ssldim"]), leaky=leaky, dtype=torch.float64)
    with torch.no_grad():
        for p, val in zip((head.meta2, head.meta2_bias, head.meta3, head.meta3_bias), _meta(fx)):
            p.copy_(torch.from_numpy(val).reshape(p.shape))
    t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    fu, fi, uv, iv = t(fx["final_user_vector"]), t(fx["final_item_vector"]), t(fx["user_vector"]), t(fx["item_vector"])
    w = head.user_weight(fu, uv)
    np.testing.assert_allclose(w.detach().numpy(), fx["user_weight"], rtol=1e-11, atol=1e-13)
    lrelu = lambda x: torch.maximum(leaky * x, x)
    loss = 0.0
    for k in range(T):
        su, si = torch.from_numpy(fx["suids%d" % k]).long(), torch.from_numpy(fx["siids%d" % k]).long()
        final_scores = lrelu(fu[su] * fi[si]).sum(-1)              # what sagnn_b200.pair_scores gathers on the GPU
        interval_scores = lrelu(uv[k][su] * iv[k][si]).sum(-1)
        loss = loss + head.hinge(w[k][su], final_scores, interval_scores)
    np.testing.assert_allclose(float(loss.detach()), float(fx["sslloss"]), rtol=1e-11)


def _seq_params(fx, dtype=np.float64):
    lay = _layout(fx)
    v = lambda j: fx["var%02d" % j].astype(dtype)
    ln = lambda o: (v(o + 1), v(o))                                   # (gamma, beta); created as beta, gamma
    return v(lay["pos"]), ln(lay["seq_ln"]), ln(lay["pos_ln"]), [_block(fx, o, dtype) for o in lay["layers"]]


@pytest.mark.parametrize("path", _DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in _DOWNSTREAM])
def test_sequence_branch_and_full_preds_match_reference_model_py_executed(path):
    """model.py:111-112,157-168 (``att_user``) and the complete ``preds`` of model.py:169-173 -- what ``ours()``
    returns next to ``sslloss`` -- oracle and product module against the executed text (padding rows included)."""
    from sagnn_b200.fusion import SequenceAttention
    fx = np.load(path)
    d, heads, leaky = int(fx["d"]), int(fx["heads"]), float(fx["leaky"])
    pos, ln_seq, ln_pos, layers = _seq_params(fx)
    fu, fi = fx["final_user_vector"], fx["final_item_vector"]
    att = fo.sequence_attention(fi, fx["sequence"], fx["mask"].astype(np.float64), pos, ln_seq, ln_pos, layers, heads, leaky)
    np.testing.assert_allclose(att, fx["att_user"], rtol=1e-10, atol=1e-11)
    preds = fo.predictions(fu, fi, att, fx["uids"], fx["iids"], fx["uLocs_seq"], leaky)
    np.testing.assert_allclose(preds, fx["preds"], rtol=1e-10, atol=1e-11)
    m = SequenceAttention(d, heads=heads, att_layers=int(fx["att_layer"]), pos_length=int(fx["pos_length"]), leaky=leaky,
                          dtype=torch.float64)
    t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    with torch.no_grad():
        m.pos_embed.copy_(t(pos))
        m.seq_ln_gamma.copy_(t(ln_seq[0])); m.seq_ln_beta.copy_(t(ln_seq[1]))
        m.pos_ln_gamma.copy_(t(ln_pos[0])); m.pos_ln_beta.copy_(t(ln_pos[1]))
        for l, blk in enumerate(layers):
            for k, val in blk.items():
                m.layer_params(l)[k].copy_(t(val))
    got = m(t(fi), torch.from_numpy(fx["sequence"]).long(), torch.from_numpy(fx["mask"]))
    np.testing.assert_allclose(got.detach().numpy(), fx["att_user"], rtol=1e-10, atol=1e-11)
    gp = m.predict(t(fu), t(fi), got, torch.from_numpy(fx["uids"]).long(), torch.from_numpy(fx["iids"]).long(),
                   torch.from_numpy(fx["uLocs_seq"]).long())
    np.testing.assert_allclose(gp.detach().numpy(), fx["preds"], rtol=1e-10, atol=1e-11)
    from sagnn_b200.fusion import prediction_hinge                   # model.py:241-244
    np.testing.assert_allclose(float(prediction_hinge(gp).detach()), float(fx["preLoss"]), rtol=1e-10)
