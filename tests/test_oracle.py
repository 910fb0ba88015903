"""CPU tests of the oracle itself: closed-form cases, golden fixtures, cross-implementation
agreement, finite differences.  (The oracle is test infrastructure; see oracle/__init__.py.)"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import c_oracle, propagate_oracle as po, tf1_mirror
from helpers import adj_lists, random_interval_mats, random_tables


def _load_mat(z):
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


# ---------------------------------------------------------------- golden: reference-generated
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "index_*.npz"))))
def test_index_construction_matches_reference_fixtures(path):
    """trans_to_lsts / transpose vs the fixtures produced by the reference's own
    DataHandler.transToLsts / transpose (tests/golden/make_golden.py)."""
    z = np.load(path)
    m = _load_mat(z)
    for norm, tag in ((False, "raw"), (True, "norm")):
        idx, dat, shp = po.trans_to_lsts(m, norm=norm)
        assert idx.dtype == np.int32 and dat.dtype == np.int32
        np.testing.assert_array_equal(idx, z[f"adj_idx_{tag}"])
        np.testing.assert_array_equal(dat, z[f"adj_data_{tag}"])
        assert shp == list(z[f"adj_shape_{tag}"])
        tidx, tdat, tshp = po.trans_to_lsts(po.transpose(m), norm=norm)
        np.testing.assert_array_equal(tidx, z[f"tp_idx_{tag}"])
        np.testing.assert_array_equal(tdat, z[f"tp_data_{tag}"])
        assert tshp == list(z[f"tp_shape_{tag}"])
    t = po.transpose(m)
    np.testing.assert_array_equal(t.indptr, z["tp_indptr"])
    np.testing.assert_array_equal(t.indices, z["tp_indices"])
    r, c = po.value_sum_degrees(m)
    np.testing.assert_array_equal(r, z["rowsum"].astype(np.int64))
    np.testing.assert_array_equal(c, z["colsum"].astype(np.int64))


def test_reference_norm_is_dead_arithmetic(golden_dir):
    """with timestamp values the int32-truncated normalisation is all zeros."""
    z = np.load(os.path.join(golden_dir, "index_ts_40x30.npz"))
    assert np.all(z["adj_data_norm"] == 0) and np.all(z["tp_data_norm"] == 0)
    assert np.all(z["adj_data_raw"] > 0)


def test_empty_matrix_fallback_edge(golden_dir):
    z = np.load(os.path.join(golden_dir, "index_empty_6x4.npz"))
    np.testing.assert_array_equal(z["adj_idx_raw"], [[0, 0]])
    np.testing.assert_array_equal(z["tp_idx_raw"], [[0, 0]])
    idx, dat, shp = po.trans_to_lsts(sp.csr_matrix((6, 4), dtype=np.intc))
    np.testing.assert_array_equal(idx, [[0, 0]])
    np.testing.assert_array_equal(dat, [0])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prop_*.npz"))))
def test_oracle_frozen_outputs(path):
    """Regression freeze of the numpy oracle (oracle-generated, not reference-generated)."""
    z = np.load(path)
    T, L = int(z["T"]), int(z["L"])
    adj = [z[f"adj{k}"] for k in range(T)]
    tp = [z[f"tp{k}"] for k in range(T)]
    out = po.propagate(adj, tp, z["uE"], z["iE"], z["gU"], z["gI"], L, float(z["leaky"]), np.float64)
    for got, name in zip(out, ("user_vec", "item_vec", "dU", "dI")):
        np.testing.assert_allclose(got, z[name], rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------- golden: the reference's model.py, executed
MODELREF = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "modelref_*.npz"))
                  if not p.endswith("modelref_errors.npz"))


@pytest.mark.parametrize("path", MODELREF)
def test_oracles_match_reference_model_py_executed(path):
    """Float parity pinned to the reference: the fixtures hold what /root/reference/model.py:80-102,
    118-134 computes when its own text is executed over numpy stand-ins for the TF ops
    (tests/golden/make_golden_model.py + tf1_shim.py), with the adjacency lists built by the
    reference's prepareModel lines 227-238; gradients are fp64 central differences of that forward.
    All three restatements (numpy, C, torch mirror) must reproduce them."""
    z = np.load(path)
    T, L, leaky = int(z["T"]), int(z["L"]), float(z["leaky"])
    adj = [z[f"adj{k}"] for k in range(T)]
    tp = [z[f"tp{k}"] for k in range(T)]
    # index construction of the fixture == the oracle's own (and therefore the device plan's, tested on GPU)
    for k in range(T):
        m = sp.csr_matrix((z[f"csr_data{k}"], z[f"csr_indices{k}"], z[f"csr_indptr{k}"]), shape=(int(z["U"]), int(z["I"])))
        np.testing.assert_array_equal(po.trans_to_lsts(m)[0], adj[k])
        np.testing.assert_array_equal(po.trans_to_lsts(po.transpose(m))[0], tp[k])
        np.testing.assert_array_equal(po.trans_to_lsts(m, norm=True)[1], z[f"adj_values{k}"])   # all zeros: F3
    for impl in (po.propagate, c_oracle.propagate):
        uv, iv, du, di = impl(adj, tp, z["uE"], z["iE"], z["gU"], z["gI"], L, leaky, np.float64)[:4]
        np.testing.assert_allclose(uv, z["user_vector"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(iv, z["item_vector"], rtol=1e-12, atol=1e-12)
        # [R,T,d] hand-off of model.py:133-134
        np.testing.assert_allclose(np.transpose(uv, (1, 0, 2)), z["user_vector_tensor"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(np.transpose(iv, (1, 0, 2)), z["item_vector_tensor"], rtol=1e-12, atol=1e-12)
        # finite differences with h = 1e-8 carry ~1e-7 of rounding noise
        assert po.relerr(du, z["dU"]) < 2e-6 and po.relerr(di, z["dI"]) < 2e-6
    # the reference computes in fp32: the fp32 oracle agrees with its fp32 run to rounding
    uv32, iv32 = po.propagate(adj, tp, z["uE"], z["iE"], z["gU"], z["gI"], L, leaky, np.float32)[:2]
    assert po.relerr(uv32, z["user_vector_f32"]) < 1e-6 and po.relerr(iv32, z["item_vector_f32"]) < 1e-6
    t64 = lambda x: torch.from_numpy(np.asarray(x, dtype=np.float64))
    t = lambda x: torch.from_numpy(np.asarray(x, dtype=np.int64))
    tu, ti, tdu, tdi = tf1_mirror.propagate([t(x) for x in adj], [t(x) for x in tp], t64(z["uE"]), t64(z["iE"]),
                                            t64(z["gU"]), t64(z["gI"]), L, leaky)
    assert po.relerr(tu.numpy(), z["user_vector"]) < 1e-12 and po.relerr(ti.numpy(), z["item_vector"]) < 1e-12
    assert po.relerr(tdu.numpy(), z["dU"]) < 2e-6 and po.relerr(tdi.numpy(), z["dI"]) < 2e-6


def test_inputs_the_reference_itself_rejects(golden_dir):
    """modelref_errors.npz: executing model.py on a one-edge interval (also the (0,0) fallback of an
    empty matrix, DataHandler.py:66-68) fails in segment_sum's shape inference (tf.squeeze made the
    ids rank 0), and rows ending more than 100 before R fail in the identity lookup (model.py:87-91).
    The oracle's strict mode raises on the second; both are inputs the reference cannot run, so the
    product's natural result there (zero rows / the fallback edge contributing) has nothing to match."""
    z = np.load(os.path.join(golden_dir, "modelref_errors.npz"))
    assert "rank 0" in str(z["one_edge"]) and "rank 0" in str(z["empty"])
    assert "is not in [0, 200)" in str(z["tail_gap_gt_100"])
    for frag in ("self.messagePropagate(embs1[-1],self.edgeDropout(self.subAdj[k]),'user')", "tf.add_n(embs0)",
                 "transToLsts(transpose(seqadj), norm=True)"):
        assert frag in str(z["executed_text"])


# ---------------------------------------------------------------- closed-form tiny cases
def test_closed_form_3x3_two_layers():
    # A = [[1,1,0],[0,0,0],[0,1,1]]   user 1 has no edges; item 0 only user 0
    A = np.array([[1, 1, 0], [0, 0, 0], [0, 1, 1]], dtype=np.float64)
    m = sp.csr_matrix(A.astype(np.intc))
    adj, tp = adj_lists([m])
    rng = np.random.default_rng(3)
    E0 = rng.standard_normal((3, 4))
    E1 = rng.standard_normal((3, 4))
    lk = 0.5
    s = lambda x: np.maximum(lk * x, x)
    e0_1 = E0 + s(A @ E1)
    e1_1 = E1 + s(A.T @ E0)
    e0_2 = e0_1 + s(A @ e1_1)
    e1_2 = e1_1 + s(A.T @ e0_1)
    uv, iv, _ = po.propagate_forward(adj, tp, E0[None], E1[None], 2, lk)
    np.testing.assert_allclose(uv[0], E0 + e0_1 + e0_2, rtol=1e-14)
    np.testing.assert_allclose(iv[0], E1 + e1_1 + e1_2, rtol=1e-14)
    # L=2 closed form of SURVEY A.1: user = 3 E0 + 2 s(Z0^0) + s(Z0^1)
    np.testing.assert_allclose(uv[0], 3 * E0 + 2 * s(A @ E1) + s(A @ e1_1), rtol=1e-13)
    # the edgeless user row is just (L+1) * its embedding
    np.testing.assert_allclose(uv[0][1], 3 * E0[1], rtol=1e-14)


def test_message_propagate_matches_dense_binary_structure():
    """Stored values (timestamps) are ignored: the operator is the binary structure."""
    m = random_interval_mats(1, 30, 20, 120, seed=5)[0]
    idx, _, _ = po.trans_to_lsts(m)
    src = np.random.default_rng(0).standard_normal((20, 8))
    B = (m.toarray() != 0).astype(np.float64)
    np.testing.assert_allclose(po.message_propagate(src, idx, 30, 0.5), np.maximum(0.5 * (B @ src), B @ src),
                               rtol=1e-13, atol=1e-13)


def test_empty_interval_uses_fallback_edge():
    """All-empty interval -> fake edge (0,0) which DOES contribute."""
    m = sp.csr_matrix((5, 4), dtype=np.intc)
    adj, tp = adj_lists([m])
    E0 = np.ones((1, 5, 2))
    E1 = 2 * np.ones((1, 4, 2))
    uv, iv, _ = po.propagate_forward(adj, tp, E0, E1, 1, 0.5)
    np.testing.assert_allclose(uv[0][0], 1 + (1 + 2))      # E0 + (E0 + s(E1[0]))
    np.testing.assert_allclose(uv[0][1], 2.0)              # untouched rows: 2 * E0
    np.testing.assert_allclose(iv[0][0], 2 + (2 + 1))


def test_strict_pad_mirrors_tf_cpu_error(golden_dir):
    z = np.load(os.path.join(golden_dir, "index_gap_300x200.npz"))
    m = _load_mat(z)
    idx, _, _ = po.trans_to_lsts(m)
    src = np.zeros((200, 4))
    assert idx[-1, 0] + 101 < 300
    with pytest.raises(IndexError):
        po.message_propagate(src, idx, 300, strict_pad=True)
    assert po.message_propagate(src, idx, 300).shape == (300, 4)   # TF-GPU behaviour: zero rows


def test_unsorted_segment_ids_rejected():
    idx = np.array([[2, 0], [1, 1]], dtype=np.int32)
    with pytest.raises(ValueError):
        po.message_propagate(np.zeros((3, 2)), idx, 3)


def test_tie_rule_zero_preactivation_takes_leaky_branch():
    """z == 0 -> gradient * leaky (TF MaximumGrad x >= y)."""
    z = np.array([[0.0, 1.0, -1.0]])
    g = np.ones_like(z)
    np.testing.assert_array_equal(po.leaky_relu_grad(z, g, 0.25), [[0.25, 1.0, 0.25]])
    m = sp.csr_matrix(np.array([[1, 0], [0, 1]], dtype=np.intc))
    adj, tp = adj_lists([m])
    E0 = np.zeros((1, 2, 3))
    E1 = np.zeros((1, 2, 3))
    G = np.ones((1, 2, 3))
    uv, iv, du, di = po.propagate(adj, tp, E0, E1, G, G, 1, 0.25)
    # n0 = G + G + A (0.25 * G) = 2.25
    np.testing.assert_allclose(du, 2.25)
    np.testing.assert_allclose(di, 2.25)


# ---------------------------------------------------------------- cross-implementation agreement
@pytest.mark.parametrize("T,U,I,d,L,nnz", [(3, 60, 45, 16, 2, 300), (2, 31, 77, 8, 3, 500), (1, 20, 20, 4, 1, 50)])
def test_numpy_c_and_torch_mirror_agree_fp64(T, U, I, d, L, nnz):
    mats = random_interval_mats(T, U, I, nnz, seed=T * 7 + L)
    adj, tp = adj_lists(mats)
    uE, iE, gU, gI = [x.astype(np.float64) for x in random_tables(T, U, I, d, seed=1)]
    a = po.propagate(adj, tp, uE, iE, gU, gI, L, 0.5, np.float64)
    b = c_oracle.propagate(adj, tp, uE, iE, gU, gI, L, 0.5, np.float64)
    t = lambda x: torch.from_numpy(x)
    c = tf1_mirror.propagate([t(x.astype(np.int64)) for x in adj], [t(x.astype(np.int64)) for x in tp],
                             t(uE), t(iE), t(gU), t(gI), L, 0.5)
    for x, y, w in zip(a, b, c):
        assert po.relerr(y, x) < 1e-13
        assert po.relerr(w.numpy(), x) < 1e-13


def test_weighted_mode_agrees():
    mats = random_interval_mats(2, 40, 30, 200, seed=9)
    adj, tp = adj_lists(mats)
    ew = [po.lightgcn_edge_weights(a, 40, 30) for a in adj]
    tew = [po.lightgcn_edge_weights(a, 30, 40) for a in tp]
    uE, iE, gU, gI = [x.astype(np.float64) for x in random_tables(2, 40, 30, 8, seed=2)]
    a = po.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float64, ew, tew)
    b = c_oracle.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float64, ew, tew)
    for x, y in zip(a, b):
        assert po.relerr(y, x) < 1e-13
    # dense check of the weighted forward, one layer
    W = np.zeros((40, 30))
    W[adj[0][:, 0], adj[0][:, 1]] = ew[0]
    z = W @ iE[0]
    uv, _, _ = po.propagate_forward(adj[:1], tp[:1], uE[:1], iE[:1], 1, 0.5, np.float64, ew[:1], tew[:1])
    np.testing.assert_allclose(uv[0], 2 * uE[0] + np.maximum(0.5 * z, z), rtol=1e-12, atol=1e-12)


def test_fp32_oracle_close_to_fp64():
    mats = random_interval_mats(2, 200, 150, 3000, seed=4)
    adj, tp = adj_lists(mats)
    uE, iE, gU, gI = random_tables(2, 200, 150, 32, seed=3)
    ref = po.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float64)
    f32 = c_oracle.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float32)
    for x, y in zip(ref, f32):
        assert po.relerr(y, x) < 1e-5


# ---------------------------------------------------------------- gradient check
def test_backward_matches_finite_differences_fp64():
    T, U, I, d, L = 2, 9, 7, 3, 2
    mats = random_interval_mats(T, U, I, 25, seed=11)
    adj, tp = adj_lists(mats)
    rng = np.random.default_rng(5)
    uE = rng.standard_normal((T, U, d))
    iE = rng.standard_normal((T, I, d))
    gU = rng.standard_normal((T, U, d))
    gI = rng.standard_normal((T, I, d))

    def loss(u, i):
        uv, iv, _ = po.propagate_forward(adj, tp, u, i, L, 0.5)
        return float((uv * gU).sum() + (iv * gI).sum())

    _, _, du, di = po.propagate(adj, tp, uE, iE, gU, gI, L, 0.5)
    eps = 1e-6
    for arr, grad, which in ((uE, du, 0), (iE, di, 1)):
        flat = arr.reshape(-1)
        for pos in rng.choice(flat.size, size=25, replace=False):
            old = flat[pos]
            flat[pos] = old + eps
            lp = loss(uE, iE)
            flat[pos] = old - eps
            lm = loss(uE, iE)
            flat[pos] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - grad.reshape(-1)[pos]) < 1e-6 * max(1.0, abs(fd)), (which, pos)


def test_pair_scores_oracle_closed_form_autograd_and_tie_rule():
    """SURVEY 8f N2 oracle (model.py:171-173,194-198): closed form on a tiny case, agreement with torch
    autograd of the same op graph away from ties, and TF's MaximumGrad tie rule at x == 0."""
    import torch
    U = np.array([[1.0, -2.0], [0.5, 4.0], [3.0, 0.0]])
    I = np.array([[2.0, 1.0], [-1.0, 0.25]])
    uids, iids = [0, 2, 1, 0], [0, 1, 1, 0]
    s, x = po.pair_scores(U, I, uids, iids, leaky=0.5)
    # sample 0: (2, -2) -> 2 + (-1) = 1; sample 1: (-3, 0) -> -1.5 + 0; sample 2: (-0.5, 1) -> -0.25 + 1
    assert np.allclose(s, [1.0, -1.5, 0.75, 1.0])
    s_lin, _ = po.pair_scores(U, I, uids, iids, activation=False)
    assert np.allclose(s_lin, [0.0, -3.0, 0.5, 0.0])
    g = np.array([1.0, 2.0, -1.0, 0.5])
    dU, dI = po.pair_scores_backward(U, I, uids, iids, g, leaky=0.5)
    # the x == 0 element (user 2, column 1) takes the leaky branch: d/du = leaky * g * i
    assert dU[2, 1] == 0.5 * 2.0 * 0.25
    # user 0 is sampled twice (g = 1 and 0.5): its gradient rows add up (unsorted_segment_sum)
    assert np.allclose(dU[0], [(1.0 + 0.5) * 1.0 * 2.0, (1.0 + 0.5) * 0.5 * 1.0])
    rng = np.random.default_rng(3)
    Ur, Ir = rng.standard_normal((30, 16)), rng.standard_normal((20, 16))
    uu, ii, gg = rng.integers(0, 30, 200), rng.integers(0, 20, 200), rng.standard_normal(200)
    for act in (True, False):
        tu = torch.from_numpy(Ur).requires_grad_(True)
        ti = torch.from_numpy(Ir).requires_grad_(True)
        xx = tu[torch.from_numpy(uu)] * ti[torch.from_numpy(ii)]
        sc = (torch.maximum(0.3 * xx, xx) if act else xx).sum(-1)
        sc.backward(torch.from_numpy(gg))
        s2, _ = po.pair_scores(Ur, Ir, uu, ii, leaky=0.3, activation=act)
        d2 = po.pair_scores_backward(Ur, Ir, uu, ii, gg, leaky=0.3, activation=act)
        assert np.allclose(s2, sc.detach().numpy(), rtol=1e-13, atol=1e-13)
        assert np.allclose(d2[0], tu.grad.numpy(), rtol=1e-12, atol=1e-12)
        assert np.allclose(d2[1], ti.grad.numpy(), rtol=1e-12, atol=1e-12)
