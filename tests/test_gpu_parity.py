"""GPU parity tests: the CUDA path (through the C ABI) vs the CPU oracle.

Integer work (CSR/CSC indices, degrees, truncated norm data) must be bit-exact; fp32
propagation must satisfy SURVEY 8(d): max|x-ref| / max|ref| <= 1e-5 per output tensor AND
allclose(rtol=1e-5, atol=1e-5*max|ref|) against the fp64 oracle.
"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import sagnn_b200 as sg
from sagnn_b200 import data_handler as dh
from oracle import c_oracle, propagate_oracle as po
from helpers import adj_lists, decode_gpu_masks, random_interval_mats, random_tables

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: 1e-5 relative, fp32


def assert_parity(got, ref, what):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    ref = np.asarray(ref, dtype=np.float64)
    scale = float(np.max(np.abs(ref)))
    err = po.relerr(got, ref)
    assert err <= TOL, "%s: max|x-ref|/max|ref| = %.3e" % (what, err)
    assert np.allclose(got, ref, rtol=TOL, atol=TOL * scale), what


def run_gpu(plan, uE, iE, gU, gI, L, leaky=0.5, want_masks=False):
    """fwd + bwd through the public autograd surface; optionally also the saved sign masks."""
    u = torch.from_numpy(uE).cuda().requires_grad_(True)
    i = torch.from_numpy(iE).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, leaky)
    masks = uv.grad_fn.masks if want_masks else None
    torch.autograd.backward([uv, iv], [torch.from_numpy(gU).cuda(), torch.from_numpy(gI).cuda()])
    torch.cuda.synchronize()
    if want_masks:
        T, U, d = uE.shape
        return uv, iv, u.grad, i.grad, decode_gpu_masks(masks, T, U, iE.shape[1], d, L)
    return uv, iv, u.grad, i.grad


def check_against_oracle(mats, d, L, leaky=0.5, seed=0, edge_weight=None, scale=1.0, tables=None,
                         max_ties=8, latdim=None, report=None, hot_rows=0):
    """Three-part parity (the derivative of LeakyReLU jumps at 0, so a pre-activation that is
    zero to within fp32 rounding may legitimately take the other branch):
      A. forward outputs vs the fp64 oracle                         <= 1e-5;
      B. saved sign masks == the oracle's except at near-ties (|z| <= 1e-5 * sum|terms|),
         and at most `max_ties` of those;
      C. backward vs the fp64 oracle GIVEN the same masks           <= 1e-5."""
    T, (U, I) = len(mats), mats[0].shape
    adj, tp = adj_lists(mats)
    uE, iE, gU, gI = tables if tables is not None else random_tables(T, U, I, d, seed=seed, scale=scale)
    ew = tew = None
    if edge_weight == "lightgcn":
        ew = [po.lightgcn_edge_weights(a, U, I) for a in adj]
        tew = [po.lightgcn_edge_weights(a, I, U) for a in tp]
    # latdim hint = d by default: plans hinted >= 128 run the v8 kernel, below that the packet-stream kernel
    plan = sg.build_plan(mats, edge_weight=edge_weight, latdim=latdim or d, hot_rows=hot_rows)
    uv, iv, du, di, gm = run_gpu(plan, uE, iE, gU, gI, L, leaky, want_masks=True)
    ref = c_oracle.propagate(adj, tp, uE, iE, gU, gI, L, leaky, np.float64, ew, tew,
                             mask_in=gm, mask_cmp=gm, tie_tol=1e-5)
    assert_parity(uv, ref[0], "user_vec")
    assert_parity(iv, ref[1], "item_vec")
    mism, far = [int(x) for x in ref[4]["stats"]]
    assert far == 0, "%d sign-mask mismatches away from ties (of %d)" % (far, mism)
    assert mism <= max_ties, "%d near-tie sign flips" % mism
    assert_parity(du, ref[2], "dU")
    assert_parity(di, ref[3], "dI")
    if report is not None:
        # also the UNCONDITIONAL backward error (oracle with its own masks) and the number of near-tie flips
        free = c_oracle.propagate(adj, tp, uE, iE, gU, gI, L, leaky, np.float64, ew, tew)
        line = ("%s: near-tie sign flips %d (far from ties %d); fwd relerr %.2e / %.2e; bwd relerr given the GPU's masks "
                "%.2e / %.2e, against the oracle's own masks %.2e / %.2e" % (
                    report, mism, far, po.relerr(uv.detach().cpu().numpy(), ref[0]), po.relerr(iv.detach().cpu().numpy(), ref[1]),
                    po.relerr(du.cpu().numpy(), ref[2]), po.relerr(di.cpu().numpy(), ref[3]),
                    po.relerr(du.cpu().numpy(), free[2]), po.relerr(di.cpu().numpy(), free[3])))
        print("\n[parity] " + line)
        os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.txt"), "a") as f:
            f.write(line + "\n")
    return plan


# ---------------------------------------------------------------- plan: bit-exact integer work
def _load_mat(z):
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "index_*.npz"))))
def test_plan_indices_match_reference_fixtures(path):
    """Device CSR/CSC vs the reference's own transToLsts / transpose outputs (golden fixtures)."""
    z = np.load(path)
    m = _load_mat(z)
    if m.dtype != np.intc:
        m = m.astype(np.intc)
    plan = sg.build_plan([m])
    np.testing.assert_array_equal(plan.adjacency_list(0, 0).cpu().numpy(), z["adj_idx_raw"])
    np.testing.assert_array_equal(plan.adjacency_list(0, 1).cpu().numpy(), z["tp_idx_raw"])
    if m.nnz:
        ptr, idx = plan.csr(0, 1)
        np.testing.assert_array_equal(ptr.cpu().numpy(), z["tp_indptr"])
        np.testing.assert_array_equal(idx.cpu().numpy(), z["tp_indices"])
        ptr, idx = plan.csr(0, 0)
        np.testing.assert_array_equal(ptr.cpu().numpy(), z["indptr"])
        np.testing.assert_array_equal(idx.cpu().numpy(), z["indices"])
        if np.issubdtype(z["data"].dtype, np.integer):
            deg_u, vs_u = plan.degrees(0, 0, value_sum=True)
            deg_i, vs_i = plan.degrees(0, 1, value_sum=True)
            np.testing.assert_array_equal(vs_u.cpu().numpy(), z["rowsum"].astype(np.int64))
            np.testing.assert_array_equal(vs_i.cpu().numpy(), z["colsum"].astype(np.int64))
            np.testing.assert_array_equal(deg_u.cpu().numpy(), np.diff(z["indptr"]))
            np.testing.assert_array_equal(deg_i.cpu().numpy(), np.diff(z["tp_indptr"]))
            np.testing.assert_array_equal(plan.norm_data(0, 0).cpu().numpy(), z["adj_data_norm"])
            np.testing.assert_array_equal(plan.norm_data(0, 1).cpu().numpy(), z["tp_data_norm"])


def test_plan_norm_data_nontrivial_values():
    """Mixed-sign stored values make the truncated normalisation non-zero (positive values alone
    can never exceed 1 before truncation): bit-exact vs the oracle, including truncation toward 0."""
    A = np.array([[1000, -990, 0, 0], [-990, 1000, 0, 0], [0, 0, 5, 0], [0, 3, 0, 7]], dtype=np.intc)
    rng = np.random.default_rng(2)
    keys = rng.choice(50 * 40, size=300, replace=False)
    R = sp.csr_matrix((rng.integers(1, 2000, size=300).astype(np.intc), (keys // 40, keys % 40)), shape=(50, 40))
    for m in (sp.csr_matrix(A), R):
        plan = sg.build_plan([m])
        for side, mat in ((0, m), (1, po.transpose(m))):
            _, ref, _ = po.trans_to_lsts(mat, norm=True)
            np.testing.assert_array_equal(plan.norm_data(0, side).cpu().numpy(), ref)
    ref = po.trans_to_lsts(sp.csr_matrix(A), norm=True)[1]
    assert ref[0] == 99 and ref[2] == -98 and ref[4] == 0


def test_plan_multi_interval_and_generator_shapes():
    g = dh.make_named("small", seed=100)
    plan = sg.build_plan(g.sub_mat)
    for k, m in enumerate(g.sub_mat):
        for side, mat in ((0, m), (1, po.transpose(m))):
            ref_idx, _, _ = po.trans_to_lsts(mat)
            np.testing.assert_array_equal(plan.adjacency_list(k, side).cpu().numpy(), ref_idx)
            rs, cs = po.value_sum_degrees(mat)
            deg, vs = plan.degrees(k, side, value_sum=True)
            np.testing.assert_array_equal(vs.cpu().numpy(), rs)
            np.testing.assert_array_equal(deg.cpu().numpy(), np.diff(sp.csr_matrix(mat).indptr))
    st = plan.stats()
    assert st["rows"] == 3 * (g.n_user + g.n_item) and st["edges_both_sides"] == 2 * sum(g.nnz)
    assert st["short_rows"] + st["long_rows"] == st["rows"]


def test_lightgcn_weights_match_oracle():
    mats = random_interval_mats(2, 70, 50, 400, seed=3)
    adj, tp = adj_lists(mats)
    plan = sg.build_plan(mats, edge_weight="lightgcn")
    for k in range(2):
        np.testing.assert_allclose(plan.weights(k, 0).cpu().numpy(),
                                   po.lightgcn_edge_weights(adj[k], 70, 50, np.float32), rtol=1e-7)
        np.testing.assert_allclose(plan.weights(k, 1).cpu().numpy(),
                                   po.lightgcn_edge_weights(tp[k], 50, 70, np.float32), rtol=1e-7)


def test_unsorted_and_out_of_range_inputs_rejected():
    row = np.array([0, 2, 1], dtype=np.int32)
    col = np.array([0, 1, 2], dtype=np.int32)
    with pytest.raises(sg.SagnnError) as e:
        sg.build_plan([(row, col)], U=4, I=4)
    assert e.value.code == 2 and "not increasing" in str(e.value)     # TF: "segment ids are not increasing"
    with pytest.raises(sg.SagnnError) as e:
        sg.build_plan([(np.array([0, 1], np.int32), np.array([0, 9], np.int32))], U=4, I=4)
    assert e.value.code == 6


def test_strict_pad_raises_like_tf_cpu(golden_dir):
    z = np.load(os.path.join(golden_dir, "index_gap_300x200.npz"))
    m = _load_mat(z)
    with pytest.raises(IndexError):
        sg.build_plan([m], strict_pad=True)
    sg.build_plan([m])   # default: zero rows, like TF-GPU


def test_cpu_tensors_are_rejected():
    plan = sg.build_plan(random_interval_mats(1, 10, 10, 20))
    with pytest.raises(RuntimeError):
        sg.propagate(plan, torch.zeros(1, 10, 64), torch.zeros(1, 10, 64), 2)
    with pytest.raises(sg.SagnnError):   # d not supported
        sg.propagate(plan, torch.zeros(1, 10, 48).cuda(), torch.zeros(1, 10, 48).cuda(), 2)


# ---------------------------------------------------------------- op-level: messagePropagate
@pytest.mark.parametrize("d", [32, 64, 128, 256])
def test_message_propagate_matches_oracle(d):
    mats = random_interval_mats(3, 150, 90, 1500, seed=d)
    adj, tp = adj_lists(mats)
    plan = sg.build_plan(mats)
    rng = np.random.default_rng(d)
    for k in (0, 2):
        src_i = rng.standard_normal((90, d)).astype(np.float32)
        src_u = rng.standard_normal((150, d)).astype(np.float32)
        got = sg.message_propagate(torch.from_numpy(src_i).cuda(), plan, k, "user", 0.5)
        assert_parity(got, po.message_propagate(src_i.astype(np.float64), adj[k], 150, 0.5), "user side")
        got = sg.message_propagate(torch.from_numpy(src_u).cuda(), plan, k, "item", 0.5)
        assert_parity(got, po.message_propagate(src_u.astype(np.float64), tp[k], 90, 0.5), "item side")


# ---------------------------------------------------------------- full path vs golden / oracle
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prop_*.npz"))))
def test_propagate_matches_golden_fixtures(path):
    z = np.load(path)
    T, U, I, L, leaky = int(z["T"]), int(z["U"]), int(z["I"]), int(z["L"]), float(z["leaky"])
    plan = sg.build_plan([z[f"adj{k}"] for k in range(T)], U=U, I=I)
    got = run_gpu(plan, z["uE"], z["iE"], z["gU"], z["gI"], L, leaky)
    for g, name in zip(got, ("user_vec", "item_vec", "dU", "dI")):
        assert_parity(g, z[name], name)


MODELREF = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "modelref_*.npz"))
                  if not p.endswith("modelref_errors.npz"))


@pytest.mark.parametrize("path", MODELREF)
@pytest.mark.parametrize("layout", ["trd", "rtd"])
def test_propagate_matches_reference_model_py_executed(path, layout):
    """CUDA path vs the outputs of the reference's own model.py text (tests/golden/make_golden_model.py):
    forward <= 1e-5 against user_vector / item_vector (and the transposed tensors of model.py:133-134 in
    the [R,T,d] layout), backward <= 1e-5 against fp64 finite differences of that forward -- no masks
    shared with the checker.  The plan is built from the scipy matrices, like prepareModel does."""
    z = np.load(path)
    T, U, I, L, leaky = int(z["T"]), int(z["U"]), int(z["I"]), int(z["L"]), float(z["leaky"])
    mats = [sp.csr_matrix((z[f"csr_data{k}"], z[f"csr_indices{k}"], z[f"csr_indptr{k}"]), shape=(U, I)) for k in range(T)]
    plan = sg.build_plan(mats)
    u = torch.from_numpy(z["uE"]).cuda().requires_grad_(True)
    i = torch.from_numpy(z["iE"]).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, leaky, layout=layout)
    gu, gi = torch.from_numpy(z["gU"]).cuda(), torch.from_numpy(z["gI"]).cuda()
    if layout == "rtd":
        assert_parity(uv, z["user_vector_tensor"], "user_vector_tensor [U,T,d]")
        assert_parity(iv, z["item_vector_tensor"], "item_vector_tensor [I,T,d]")
        gu, gi = gu.transpose(0, 1).contiguous(), gi.transpose(0, 1).contiguous()
    else:
        assert_parity(uv, z["user_vector"], "user_vector [T,U,d]")
        assert_parity(iv, z["item_vector"], "item_vector [T,I,d]")
    torch.autograd.backward([uv, iv], [gu, gi])
    assert_parity(u.grad, z["dU"], "dU vs finite differences of the reference forward")
    assert_parity(i.grad, z["dI"], "dI vs finite differences of the reference forward")
    for k in range(T):   # adjacency lists: the reference's transToLsts / transpose outputs
        ptr, idx = plan.csr(k, 0)
        np.testing.assert_array_equal(idx.cpu().numpy(), z[f"adj{k}"][:, 1])
        ptr, idx = plan.csr(k, 1)
        np.testing.assert_array_equal(idx.cpu().numpy(), z[f"tp{k}"][:, 1])


@pytest.mark.parametrize("d,L", [(32, 1), (64, 2), (64, 3), (128, 3), (128, 4), (256, 2)])
def test_propagate_small_random(d, L):
    check_against_oracle(random_interval_mats(3, 120, 80, 900, seed=d + L), d, L, seed=L)


@pytest.mark.parametrize("d,L,hint", [(128, 3, 64), (256, 2, 64), (64, 2, 128), (32, 1, 128)])
def test_propagate_other_kernel_than_the_hint_would_pick(d, L, hint):
    """A plan hinted below 128 carries the packed task stream (v10 kernel), from 128 on the v8 task records:
    both kernels must serve every latdim."""
    check_against_oracle(random_interval_mats(3, 120, 80, 900, seed=d + L), d, L, seed=L, latdim=hint)


@pytest.mark.parametrize("d,L,hot,weights", [(64, 2, 32, None), (64, 3, 500, None), (32, 2, 100, "lightgcn"),
                                              (64, 2, 200, "lightgcn"), (128, 2, 64, None)])
def test_propagate_hot_rows_staged_in_shared_memory(d, L, hot, weights):
    """north_star's "TMA/shared-memory staging of hot item rows" (``sagnn_plan_set_hot_rows``, off by default):
    the highest-degree source rows are read from their shared-memory copies, hot edges first inside every task --
    power-law graphs with hubs, long (sliced) rows, more hot slots asked for than fit / than rows exist, and a
    plan run at another latdim than its hint (slots that do not fit are read through their row ids)."""
    rng = np.random.default_rng(hot)
    U, I = 600, 260
    mats = []
    for k in range(3):
        pu = 1.0 / np.arange(1, U + 1) ** 0.8; pi = 1.0 / np.arange(1, I + 1)
        r = rng.choice(U, 9000, p=pu / pu.sum()); c = rng.choice(I, 9000, p=pi / pi.sum())
        A = sp.coo_matrix((np.ones(9000, np.intc), (rng.permutation(U)[r], rng.permutation(I)[c])), shape=(U, I)).tocsr()
        A.data[:] = 1
        mats.append(A)
    plan = check_against_oracle(mats, d, L, seed=hot, scale=0.1, edge_weight=weights, latdim=64 if d == 128 else d,
                                hot_rows=hot)
    st = plan.stats()
    assert 0 < st["hot_rows"] <= hot and st["long_rows"] > 0


def test_propagate_long_rows_chunked_reduction():
    """Rows far above the 64-edge chunk (dense-ish block + one hub item / hub user)."""
    U, I = 700, 300
    rng = np.random.default_rng(1)
    A = (rng.random((U, I)) < 0.05)
    A[:, 7] = True            # hub item: 700 users  -> 11 chunks on the item side
    A[13, :] = True           # hub user: 300 items
    mats = [sp.csr_matrix(A.astype(np.intc)), sp.csr_matrix((rng.random((U, I)) < 0.3).astype(np.intc))]
    plan = check_against_oracle(mats, 64, 2, seed=3, scale=0.1)
    st = plan.stats()
    assert st["long_rows"] > 0 and st["chunks"] > st["long_rows"] and st["max_degree"] == 700
    check_against_oracle(mats, 128, 3, seed=4, scale=0.05)


def test_propagate_empty_interval_and_empty_rows():
    mats = random_interval_mats(3, 40, 30, 60, seed=8)
    mats[1] = sp.csr_matrix((40, 30), dtype=np.intc)     # all-empty interval -> fallback edge (0,0)
    check_against_oracle(mats, 64, 2, seed=1)


def test_more_intervals_than_half_the_sms():
    """T = 80 intervals (160 segments > 148 SMs): the layer runs in waves of whole intervals."""
    mats = random_interval_mats(80, 40, 30, 90, seed=80)
    check_against_oracle(mats, 64, 2, seed=8)


def test_propagate_tie_rule_zero_inputs():
    """Zero embeddings make every pre-activation exactly 0: sigma'(0) must be `leaky`."""
    mats = random_interval_mats(1, 20, 20, 60, seed=2)
    adj, tp = adj_lists(mats)
    plan = sg.build_plan(mats)
    z = np.zeros((1, 20, 64), np.float32)
    g = np.ones((1, 20, 64), np.float32)
    ref = po.propagate(adj, tp, z, z, g, g, 2, 0.25, np.float64)
    got = run_gpu(plan, z, z, g, g, 2, 0.25)
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=1e-6)


def test_propagate_weighted_lightgcn_mode():
    check_against_oracle(random_interval_mats(2, 200, 150, 2500, seed=6), 64, 2, seed=2, edge_weight="lightgcn")


def test_propagate_custom_weights():
    mats = random_interval_mats(2, 60, 50, 500, seed=12)
    adj, tp = adj_lists(mats)
    rng = np.random.default_rng(0)
    ew = [rng.random(len(a)).astype(np.float32) for a in adj]
    # weights of the transposed lists: same edges, transposed order
    tew = []
    for k, m in enumerate(mats):
        w = sp.csr_matrix((ew[k], (adj[k][:, 0], adj[k][:, 1])), shape=m.shape)
        tew.append(np.asarray(sp.csr_matrix(w.T).tocoo().data, dtype=np.float32))
    uE, iE, gU, gI = random_tables(2, 60, 50, 64, seed=5)
    ref = po.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float64, ew, tew)
    plan = sg.build_plan(mats, edge_weight=ew)
    got = run_gpu(plan, uE, iE, gU, gI, 2)
    for g, r, name in zip(got, ref, ("user_vec", "item_vec", "dU", "dI")):
        assert_parity(g, r, name)


def test_repeated_calls_are_bitwise_deterministic():
    g = dh.make_named("small", seed=5)
    plan = sg.build_plan(g.sub_mat)
    uE, iE, gU, gI = random_tables(3, g.n_user, g.n_item, 64, seed=9)
    a = run_gpu(plan, uE, iE, gU, gI, 2)
    b = run_gpu(plan, uE, iE, gU, gI, 2)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_windowed_schedule_order_is_bitwise_equal(monkeypatch):
    """SAGNN_SORT_WINDOW (read at plan finalize) only permutes the task schedule: degree classes sorted inside
    windows of consecutive row ids instead of over the whole table.  Every row is still summed by one task in its own
    edge order, so outputs and gradients must not change by a bit (both kernels)."""
    g = dh.make_named("small", seed=5)
    uE, iE, gU, gI = random_tables(3, g.n_user, g.n_item, 64, seed=9)
    for latdim in (64, 128):                             # 64: packet kernel schedule, 128: v8 schedule
        monkeypatch.delenv("SAGNN_SORT_WINDOW", raising=False)
        a = run_gpu(sg.build_plan(g.sub_mat, latdim=latdim), uE, iE, gU, gI, 2)
        for w in ("64", "1000"):
            monkeypatch.setenv("SAGNN_SORT_WINDOW", w)
            b = run_gpu(sg.build_plan(g.sub_mat, latdim=latdim), uE, iE, gU, gI, 2)
            for x, y in zip(a, b):
                assert torch.equal(x, y)
    monkeypatch.delenv("SAGNN_SORT_WINDOW", raising=False)


def test_forward_only_and_partial_grad():
    mats = random_interval_mats(2, 50, 40, 300, seed=4)
    adj, tp = adj_lists(mats)
    plan = sg.build_plan(mats)
    uE, iE, gU, gI = random_tables(2, 50, 40, 64, seed=6)
    with torch.no_grad():
        uv, iv = sg.propagate(plan, torch.from_numpy(uE).cuda(), torch.from_numpy(iE).cuda(), 2)
    ref = po.propagate(adj, tp, uE, iE, gU, np.zeros_like(gI), 2, 0.5, np.float64)
    assert_parity(uv, ref[0], "user_vec")
    assert_parity(iv, ref[1], "item_vec")
    u = torch.from_numpy(uE).cuda().requires_grad_(True)
    i = torch.from_numpy(iE).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, 2)
    (uv * torch.from_numpy(gU).cuda()).sum().backward()     # item output unused -> zero upstream
    assert_parity(u.grad, ref[2], "dU")
    assert_parity(i.grad, ref[3], "dI")


def test_host_entry_point_matches_oracle():
    mats = random_interval_mats(2, 90, 70, 600, seed=14)
    adj, tp = adj_lists(mats)
    plan = sg.build_plan(mats)
    uE, iE, gU, gI = random_tables(2, 90, 70, 64, seed=7)
    ref = po.propagate(adj, tp, uE, iE, gU, gI, 2, 0.5, np.float64)
    outs = [np.empty_like(x) for x in (uE, iE, uE, iE)]
    sg.propagate_host(plan, uE, iE, gU, gI, outs[0], outs[1], outs[2], outs[3], 2, 0.5)
    for g, r, name in zip(outs, ref, ("user_vec", "item_vec", "dU", "dI")):
        assert_parity(g, r, name)
    fo = [np.empty_like(uE), np.empty_like(iE)]
    sg.propagate_host(plan, uE, iE, None, None, fo[0], fo[1], None, None, 2, 0.5)
    np.testing.assert_array_equal(fo[0], outs[0])
    # split entry points (forward now, backward later) give the same bits
    so = [np.empty_like(x) for x in (uE, iE, uE, iE)]
    with pytest.raises(sg.SagnnError):
        sg.host_backward(plan, gU, gI, so[2], so[3], 2, 0.5)        # no forward with masks yet
    sg.host_forward(plan, uE, iE, so[0], so[1], 2, 0.5, keep_masks=True)
    sg.host_backward(plan, gU, gI, so[2], so[3], 2, 0.5)
    for a, b in zip(so, outs):
        np.testing.assert_array_equal(a, b)


def test_interval_launches_equal_full_launch_bitwise():
    """Per-interval launches (used to pipeline copies with compute) give the same bits as the
    all-interval launch, and the CUDA-graph replay the same bits as direct launches."""
    from sagnn_b200.step import PropagationStep
    g = dh.make_named("small", seed=11)
    plan = sg.build_plan(g.sub_mat)
    for L, d in ((2, 64), (3, 128)):
        st = PropagationStep(plan, L, d)
        gen = torch.Generator(device="cuda").manual_seed(L)
        for t in (st.u_embed, st.i_embed, st.g_user, st.g_item):
            t.normal_(generator=gen)
        st.run()
        torch.cuda.synchronize()
        ref = [t.clone() for t in (st.user_out, st.item_out, st.d_u, st.d_i)]
        for t in (st.user_out, st.item_out, st.d_u, st.d_i):
            t.zero_()
        for k in range(plan.T):
            st.forward_interval(k)
        for k in reversed(range(plan.T)):
            st.backward_interval(k)
        torch.cuda.synchronize()
        for a, b in zip(ref, (st.user_out, st.item_out, st.d_u, st.d_i)):
            assert torch.equal(a, b)
        split = st.calibrate(rounds=1)                    # re-dealing the CTAs must not change a bit
        assert sum(split) == plan.stats()["sms"] and min(split) >= 1
        st.capture()
        for t in (st.user_out, st.item_out, st.d_u, st.d_i):
            t.zero_()
        st.replay()
        torch.cuda.synchronize()
        for a, b in zip(ref, (st.user_out, st.item_out, st.d_u, st.d_i)):
            assert torch.equal(a, b)


@pytest.mark.parametrize("d,L", [(32, 1), (64, 2), (64, 3), (128, 3), (256, 2)])
def test_rtd_layout_equals_transposed_default_bitwise(d, L):
    """SAGNN_LAYOUT_RTD (the [R,T,d] hand-off of model.py:133-134 fused into the epilogue, and the
    backward reading its upstream in that layout): same bits as the default layout, transposed;
    and parity against the oracle's outputs transposed."""
    g = dh.make_named("small", seed=21)
    plan = sg.build_plan(g.sub_mat)
    T, U, I = plan.T, g.n_user, g.n_item
    uE, iE, gU, gI = random_tables(T, U, I, d, seed=31)
    ref = run_gpu(plan, uE, iE, gU, gI, L)
    u = torch.from_numpy(uE).cuda().requires_grad_(True)
    i = torch.from_numpy(iE).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, 0.5, layout="rtd")
    assert uv.shape == (U, T, d) and iv.shape == (I, T, d) and uv.is_contiguous() and iv.is_contiguous()
    gUt = torch.from_numpy(gU).cuda().transpose(0, 1).contiguous()
    gIt = torch.from_numpy(gI).cuda().transpose(0, 1).contiguous()
    torch.autograd.backward([uv, iv], [gUt, gIt])
    torch.cuda.synchronize()
    assert torch.equal(uv.transpose(0, 1), ref[0]) and torch.equal(iv.transpose(0, 1), ref[1])
    assert torch.equal(u.grad, ref[2]) and torch.equal(i.grad, ref[3])
    adj, tp = adj_lists(g.sub_mat)
    oref = po.propagate(adj, tp, uE, iE, gU, gI, L, 0.5, np.float64)
    assert_parity(uv, np.transpose(oref[0], (1, 0, 2)), "user_vector_tensor [U,T,d]")
    assert_parity(iv, np.transpose(oref[1], (1, 0, 2)), "item_vector_tensor [I,T,d]")


def test_rtd_layout_step_graph_and_sliced_rows():
    """PropagationStep(layout='rtd'): direct launches == CUDA-graph replay == default layout, on a
    shape with sliced long rows; unknown flags are rejected."""
    import ctypes
    from sagnn_b200 import _lib
    from sagnn_b200.step import PropagationStep
    from sagnn_b200.propagate import _ptr, _stream_ptr
    mats = random_interval_mats(2, 300, 40, 6000, seed=8)          # item rows of ~150 edges: slices
    plan = sg.build_plan(mats)
    a, b = PropagationStep(plan, 2, 64), PropagationStep(plan, 2, 64, layout="rtd")
    gen = torch.Generator(device="cuda").manual_seed(3)
    for t in (a.u_embed, a.i_embed, a.g_user, a.g_item):
        t.normal_(generator=gen)
    b.u_embed.copy_(a.u_embed); b.i_embed.copy_(a.i_embed)
    b.g_user.copy_(a.g_user.transpose(0, 1)); b.g_item.copy_(a.g_item.transpose(0, 1))
    a.run(); b.run()
    torch.cuda.synchronize()
    def same():
        return (torch.equal(b.user_out.transpose(0, 1), a.user_out) and torch.equal(b.item_out.transpose(0, 1), a.item_out)
                and torch.equal(b.d_u, a.d_u) and torch.equal(b.d_i, a.d_i))
    assert same()
    b.capture()
    for t in (b.user_out, b.item_out, b.d_u, b.d_i):
        t.zero_()
    b.replay()
    torch.cuda.synchronize()
    assert same()
    rc = _lib.load_library().sagnn_propagate_fwd_ex(
        plan.handle, _ptr(b.u_embed), _ptr(b.i_embed), _ptr(b.user_out), _ptr(b.item_out), 2, 64, 0.5,
        _ptr(b.masks), _ptr(b.ws), b.ws.numel(), 0x80, _stream_ptr(plan.device))
    assert rc == 1   # SAGNN_INVALID_ARG


@pytest.mark.parametrize("W,L,d", [(1, 2, 64), (3, 2, 64), (4, 3, 128), (2, 1, 32)])
def test_fused_scatter_to_row_sharded_consumer(W, L, d):
    """sagnn_propagate_fwd_scatter: the last layer writes every row of the layer sums into the receive
    buffer of the rank owning its row block ([source rank, blk, T, d]).  W receive buffers on ONE GPU
    stand in for the peers' (peer memory is just another mapped pointer); this rank plays source
    rank `me`.  Same bits as the plain forward; rows of other source ranks and pad rows untouched."""
    from sagnn_b200.step import PropagationStep
    g = dh.make_named("small", seed=23)
    plan = sg.build_plan(g.sub_mat, latdim=d)
    T, U, I = plan.T, g.n_user, g.n_item
    a, b = PropagationStep(plan, L, d), PropagationStep(plan, L, d, layout="rtd")
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in (a.u_embed, a.i_embed):
        t.normal_(generator=gen)
    b.u_embed.copy_(a.u_embed); b.i_embed.copy_(a.i_embed)
    a.forward()
    me = W - 1
    bu, bi = -(-U // W), -(-I // W)
    rcv_u = [torch.full((W, bu, T, d), -7.0, device="cuda") for _ in range(W)]
    rcv_i = [torch.full((W, bi, T, d), -7.0, device="cuda") for _ in range(W)]
    b.set_scatter(W, me, [t.data_ptr() for t in rcv_u], [t.data_ptr() for t in rcv_i])
    b.forward()
    torch.cuda.synchronize()
    for rcv, blk, rows, ref in ((rcv_u, bu, U, a.user_out), (rcv_i, bi, I, a.item_out)):
        for r in range(W):
            lo, hi = r * blk, min((r + 1) * blk, rows)
            assert torch.equal(rcv[r][me, :hi - lo], ref[:, lo:hi].transpose(0, 1))      # [blk, T, d]
            assert (rcv[r][me, hi - lo:] == -7.0).all()                                  # pad rows untouched
            for other in range(W):
                if other != me:
                    assert (rcv[r][other] == -7.0).all()
    b.set_scatter(0, 0, None, None)                      # and the backward is unaffected by the hand-off
    b.g_user.normal_(generator=gen); b.g_item.normal_(generator=gen)
    a.g_user.copy_(b.g_user.transpose(0, 1)); a.g_item.copy_(b.g_item.transpose(0, 1))
    a.backward(); b.backward()
    torch.cuda.synchronize()
    assert torch.equal(a.d_u, b.d_u) and torch.equal(a.d_i, b.d_i)


@pytest.mark.parametrize("W,L,d", [(2, 2, 64), (3, 2, 64), (2, 1, 32)])
def test_chain_propagation_scatter_fusion_backward_on_virtual_ranks(W, L, d):
    """The whole data-parallel chain the fused hand-off is designed for, with W ranks played on
    one GPU: rank r owns the intervals LPT gives it and runs its forward with sagnn_propagate_fwd_scatter into
    the W receive buffers ([source rank, blk, T_local, d]); consumer rank j assembles row block j of ALL
    intervals from its slabs (slabs_to_rtd), runs the interval fusion (LSTM over T -> layer norm -> MHSA -> mean,
    model.py:135-155) and back-propagates a loss; the dense [blk, T, d] upstream goes back to the interval owners,
    whose sagnn_propagate_bwd_ex ([R,T,d] upstream) yields the embedding gradients.  Must equal the single-GPU
    chain propagate(layout="rtd") -> fusion -> autograd."""
    from sagnn_b200.step import PropagationStep
    from sagnn_b200.fusion import IntervalFusion, slabs_to_rtd
    from sagnn_b200.dist import assign_intervals
    mats = random_interval_mats(4, 210, 130, 1800, seed=41)
    T, U, I = 4, 210, 130
    uE, iE, _, _ = random_tables(T, U, I, d, seed=42, scale=0.3)
    fusion = IntervalFusion(d, heads=16 if d % 16 == 0 else 8, device="cuda", seed=5)
    wu = torch.randn((U, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    wi = torch.randn((I, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    # ---- single GPU
    plan = sg.build_plan(mats, latdim=d)
    u = torch.from_numpy(uE).cuda().requires_grad_(True)
    i = torch.from_numpy(iE).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, 0.5, layout="rtd")
    fu, fi = fusion(uv, iv)
    ((fu * wu).sum() + (fi * wi).sum()).backward()
    ref_fu, ref_fi, ref_du, ref_di = fu.detach(), fi.detach(), u.grad.clone(), i.grad.clone()
    fusion.zero_grad()
    # ---- W ranks: interval owners (propagation) x row-block owners (fusion)
    owners = assign_intervals([m.nnz for m in mats], W)
    mine = [[k for k in range(T) if owners[k] == r] for r in range(W)]
    tl = max(len(x) for x in mine)
    bu, bi = -(-U // W), -(-I // W)
    rcv_u = [torch.zeros((W, bu, tl, d), device="cuda") for _ in range(W)]
    rcv_i = [torch.zeros((W, bi, tl, d), device="cuda") for _ in range(W)]
    steps = []
    for r in range(W):
        st = None
        if mine[r]:
            st = PropagationStep(sg.build_plan([mats[k] for k in mine[r]], latdim=d), L, d, layout="rtd")
            st.u_embed.copy_(torch.from_numpy(uE[mine[r]])); st.i_embed.copy_(torch.from_numpy(iE[mine[r]]))
            # this rank's slabs have T_local = len(mine[r]) intervals: point the scatter at views of that depth
            vu = [t[:, :, :len(mine[r])].contiguous() for t in rcv_u]
            vi = [t[:, :, :len(mine[r])].contiguous() for t in rcv_i]
            st.set_scatter(W, r, [t.data_ptr() for t in vu], [t.data_ptr() for t in vi])
            st.forward()
            torch.cuda.synchronize()
            for j in range(W):
                rcv_u[j][r, :, :len(mine[r])] = vu[j][r]
                rcv_i[j][r, :, :len(mine[r])] = vi[j][r]
        steps.append(st)
    g_back_u = [torch.zeros((U, len(mine[r]), d), device="cuda") for r in range(W)]
    g_back_i = [torch.zeros((I, len(mine[r]), d), device="cuda") for r in range(W)]
    for j in range(W):                                    # consumer rank j: row block j of every interval
        lo_u, hi_u, lo_i, hi_i = j * bu, min((j + 1) * bu, U), j * bi, min((j + 1) * bi, I)
        xu = slabs_to_rtd(rcv_u[j], owners, hi_u - lo_u).clone().requires_grad_(True)
        xi = slabs_to_rtd(rcv_i[j], owners, hi_i - lo_i).clone().requires_grad_(True)
        fu, fi = fusion(xu, xi)
        assert_parity(fu, ref_fu[lo_u:hi_u].cpu().numpy(), "fused user vector, row block %d" % j)
        assert_parity(fi, ref_fi[lo_i:hi_i].cpu().numpy(), "fused item vector, row block %d" % j)
        ((fu * wu[lo_u:hi_u]).sum() + (fi * wi[lo_i:hi_i]).sum()).backward()
        for r in range(W):                                # reverse hand-off: upstream of interval k goes to its owner
            for jj, k in enumerate(mine[r]):
                g_back_u[r][lo_u:hi_u, jj] = xu.grad[:, k]
                g_back_i[r][lo_i:hi_i, jj] = xi.grad[:, k]
    for r in range(W):
        if steps[r] is None:
            continue
        st = steps[r]
        st.set_scatter(0, 0, None, None)
        st.g_user.copy_(g_back_u[r]); st.g_item.copy_(g_back_i[r])
        st.backward()
        torch.cuda.synchronize()
        assert_parity(st.d_u, ref_du[mine[r]].cpu().numpy(), "dU of rank %d's intervals" % r)
        assert_parity(st.d_i, ref_di[mine[r]].cpu().numpy(), "dI of rank %d's intervals" % r)


# ---------------------------------------------------------------- device-side trans_sub
def _coo_of(m):
    m = sp.csr_matrix(m)
    m.sort_indices()
    return np.repeat(np.arange(m.shape[0]), np.diff(m.indptr)), m.indices, m.data


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "trnmat_*.npz"))))
def test_bucket_events_matches_reference_notebook_fixtures(path):
    """sagnn_bucket_events vs the outputs of the reference notebook's own `trans_sub` cell
    (tests/golden/make_golden_trnmat.py): every interval's adjacency list and stored timestamps, bit-exact."""
    z = np.load(path)
    U, I, T = int(z["U"]), int(z["I"]), int(z["T"])
    lists = sg.bucket_events(z["u"], z["i"], z["t"], U, I, T)
    assert len(lists) == T
    for k, (row, col, val) in enumerate(lists):
        ind = z["sub%d_indptr" % k]
        assert np.array_equal(row.cpu().numpy(), np.repeat(np.arange(U), np.diff(ind)))
        assert np.array_equal(col.cpu().numpy(), z["sub%d_indices" % k])
        assert np.array_equal(val.cpu().numpy(), z["sub%d_data" % k])


def test_bucket_events_large_random_equals_host_mirror_and_feeds_the_plan():
    rng = np.random.default_rng(41)
    U, I, T, n = 3000, 2000, 6, 300000
    u = np.sort(rng.integers(0, U, n))                       # the notebook visits users in ascending order
    i = rng.integers(0, 60, n) * 33 % I                      # few items per user: many repeated pairs
    t = rng.integers(1388534400, 1406073600, n)
    t[:50] = t[0]                                            # equal timestamps, and the maximum lands in the clamp
    subs, _ = dh.trans_sub(u, i, t, U, I, T, *dh.trans(u, i, t, U, I)[1:])
    lists = sg.bucket_events(u, i, t, U, I, T)
    assert sum(len(l[0]) for l in lists) < n                 # duplicates were dropped
    for k in range(T):
        r, c, v = _coo_of(subs[k])
        assert np.array_equal(lists[k][0].cpu().numpy(), r) and np.array_equal(lists[k][1].cpu().numpy(), c)
        assert np.array_equal(lists[k][2].cpu().numpy(), v)
    plan, ref = sg.build_plan(lists, U, I), sg.build_plan(subs)
    for k in range(T):
        for side in (0, 1):
            assert torch.equal(plan.adjacency_list(k, side), ref.adjacency_list(k, side))
            assert torch.equal(plan.degrees(k, side, value_sum=True)[1], ref.degrees(k, side, value_sum=True)[1])
    # an interval nobody falls into (minn far below the data) becomes the reference's fallback edge
    sparse = sg.bucket_events(u, i, t, U, I, 3, minn=int(t.min()) - 10 * int(t.max() - t.min()))
    assert len(sparse[0][0]) == 0 and sg.build_plan(sparse, U, I).nnz[0] == 1
    with pytest.raises(sg.SagnnError):
        sg.bucket_events(np.array([U, 0]), np.array([0, 0]), np.array([1400000000, 1400000100]), U, I, T)   # id out of range
    with pytest.raises(sg.SagnnError):
        sg.bucket_events(np.array([0]), np.array([0]), np.array([5]), U, I, T, minn=5, maxx=5)


# ---------------------------------------------------------------- device-side sampleSslBatch
def test_sample_ssl_batch_contract():
    """sagnn_sample_ssl_batch vs the contract of Recommender.sampleSslBatch (model.py:304-339): per
    interval and batch user 2*min(sslNum, |posset|//2) samples, all of them items of that user in that
    interval, positives / negatives interleaved, uLocs / uLocs_seq laid out like the reference's, users
    with fewer than two items silent; seed-deterministic; draws uniform over posset."""
    g = dh.make_named("small", seed=29)
    plan = sg.build_plan(g.sub_mat)
    U, T, ssl = g.n_user, plan.T, 5
    rng = np.random.default_rng(1)
    bat = rng.permutation(U)[:512].astype(np.int32)
    out = plan.sample_ssl_batch(bat, ssl, seed=7)
    again = plan.sample_ssl_batch(bat, ssl, seed=7)
    other = plan.sample_ssl_batch(bat, ssl, seed=8)
    assert len(out) == T
    for k in range(T):
        m = sp.csr_matrix(g.sub_mat[k])
        deg = np.diff(m.indptr)[bat]
        want = 2 * np.minimum(ssl, deg // 2)
        u, i, sq = (np.array(x.cpu().numpy(), dtype=np.int64) for x in out[k])   # own, writeable copies (scipy indexing)
        assert len(u) == len(i) == len(sq) == int(want.sum())
        assert np.array_equal(sq, np.repeat(np.arange(len(bat)), want))          # batch order, 2*s entries each
        assert np.array_equal(u, bat[sq])
        assert np.asarray(m[u, i]).ravel().astype(bool).all()                    # every sample is an item of that user
        for a, b in zip(out[k], again[k]):
            assert torch.equal(a, b)
        assert not torch.equal(out[k][1], other[k][1])
        for a, b in zip(out[k], plan.sample_ssl_interval(k, bat, ssl, seed=7)):  # the per-interval entry point draws the same
            assert torch.equal(a, b)
    # uniformity: one user with many items, many draws -> every item drawn, counts within 5 sigma
    k, m = 0, sp.csr_matrix(g.sub_mat[0])
    hub = int(np.argmax(np.diff(m.indptr)))
    d = int(np.diff(m.indptr)[hub])
    assert d >= 8
    draws = plan.sample_ssl_batch(np.full(4000, hub, dtype=np.int32), d // 2, seed=3)[0][1].cpu().numpy()
    assert len(draws) == 4000 * 2 * (d // 2)
    items, counts = np.unique(draws, return_counts=True)
    assert np.array_equal(items, m.indices[m.indptr[hub]:m.indptr[hub + 1]])
    mean = len(draws) / d
    assert np.all(np.abs(counts - mean) < 5 * np.sqrt(mean))
    empty = plan.sample_ssl_batch(np.empty(0, dtype=np.int32), ssl)
    assert all(x[0].numel() == 0 for x in empty)
    with pytest.raises(sg.SagnnError):
        plan.sample_ssl_batch(np.array([U], dtype=np.int32), ssl)


# ---------------------------------------------------------------- sampled pair scores
def test_sample_train_batch_contract():
    """sagnn_sample_train_batch = Recommender.sampleTrainBatch + negSamp (model.py:252-302, DataHandler.py:28-41) on the
    device: layout (positives first, then negatives), counts min(train_sample_num, len(seq)-1), the positive
    posset[-choose] with choose in the reference's randint range, negatives without any training interaction and
    different from the user's last and held-out item, right-aligned history / mask, zero padding rows, determinism."""
    rng = np.random.default_rng(12)
    U, I, T = 400, 300, 3
    mats = random_interval_mats(T, U, I, 2500, seed=13)
    plan = sg.build_plan(mats)
    label = sum((m != 0).astype(np.int8) for m in mats).toarray() > 0            # trnMat structure = union of the intervals
    seqs = []
    for u in range(U):
        n = int(rng.integers(0, 30)) if u % 17 else 0                            # some users with empty / tiny sequences
        seqs.append([int(x) for x in rng.integers(0, I, size=n)])
    seqs[5] = [int(x) for x in rng.integers(0, I, size=260)]                     # longer than pos_length
    tst = [None if u % 5 == 0 else int(rng.integers(0, I)) for u in range(U)]
    bat = rng.permutation(U)[:96].astype(np.int32)
    bat[0] = 5
    tsn, pred, plen, pad = 7, 5, 200, 128
    out = plan.sample_train_batch(bat, seqs, tst, tsn, pred, plen, batch_pad=pad, seed=99)
    uL, iL, seq, mask, uS, choose = [t.cpu().numpy() for t in out]
    again = [t.cpu().numpy() for t in plan.sample_train_batch(bat, seqs, tst, tsn, pred, plen, batch_pad=pad, seed=99)]
    assert all(np.array_equal(a, b) for a, b in zip((uL, iL, seq, mask, uS, choose), again))
    other = plan.sample_train_batch(bat, seqs, tst, tsn, pred, plen, batch_pad=pad, seed=100)[1].cpu().numpy()
    assert not np.array_equal(other, iL)
    counts = [min(tsn, max(len(seqs[u]) - 1, 0)) for u in bat]
    half = sum(counts)
    assert len(uL) == len(iL) == len(uS) == 2 * half
    assert seq.shape == (pad, plen) and mask.shape == (pad, plen)
    cur = 0
    for b, u in enumerate(bat):
        s_, n = seqs[u], counts[b]
        posset = s_[:-1]
        hi = max(min(pred + 1, len(posset) - 3), 1)
        assert 1 <= choose[b] <= hi
        for j in range(n):
            for o in (cur + j, half + cur + j):
                assert uL[o] == u and uS[o] == b
            assert iL[cur + j] == posset[-choose[b]]
            neg = iL[half + cur + j]
            assert 0 <= neg < I and not label[u, neg] and neg != s_[-1] and neg != (tst[u] if tst[u] is not None else -1)
        cur += n
        hist = posset[:len(posset) - choose[b]] if n else posset[:max(len(posset) - 1, 0)]
        keep = hist[-plen:]
        want = np.zeros(plen, np.int64); wm = np.zeros(plen)
        if keep:
            want[-len(keep):] = keep; wm[-len(keep):] = 1
        np.testing.assert_array_equal(seq[b], want)
        np.testing.assert_array_equal(mask[b], wm)
    assert not seq[len(bat):].any() and not mask[len(bat):].any()
    negs = iL[half:]
    assert len(np.unique(negs)) > I // 3                                           # spread over the item range


@pytest.mark.parametrize("d,layout,act", [(64, "trd", "leakyRelu"), (64, "rtd", "leakyRelu"), (128, "trd", None),
                                          (32, "rtd", None), (256, "trd", "leakyRelu")])
def test_pair_scores_match_oracle(d, layout, act):
    """sagnn_pair_scores_fwd / _bwd (model.py:171-173,194-198) vs the fp64 oracle, on the outputs'
    two layouts; users repeat in the samples (like suids), so gradient rows collide."""
    T, U, I, n, k = 3, 70, 50, 4000, 1
    rng = np.random.default_rng(d)
    uv, iv = rng.standard_normal((T, U, d)).astype(np.float32), rng.standard_normal((T, I, d)).astype(np.float32)
    uids, iids = rng.integers(0, U, n).astype(np.int32), rng.integers(0, I, n).astype(np.int32)
    g = rng.standard_normal(n).astype(np.float32)
    tu, ti = torch.from_numpy(uv).cuda(), torch.from_numpy(iv).cuda()
    if layout == "rtd":
        tu, ti = tu.transpose(0, 1).contiguous(), ti.transpose(0, 1).contiguous()
    tu.requires_grad_(True); ti.requires_grad_(True)
    s = sg.pair_scores(tu, ti, k, torch.from_numpy(uids).cuda(), torch.from_numpy(iids).cuda(), activation=act,
                       leaky=0.5, layout=layout)
    s.backward(torch.from_numpy(g).cuda())
    torch.cuda.synchronize()
    ref_s, _ = po.pair_scores(uv[k].astype(np.float64), iv[k].astype(np.float64), uids, iids, 0.5, act is not None)
    ref_du, ref_di = po.pair_scores_backward(uv[k].astype(np.float64), iv[k].astype(np.float64), uids, iids,
                                             g.astype(np.float64), 0.5, act is not None)
    assert_parity(s, ref_s, "scores")
    du, di = (tu.grad.transpose(0, 1), ti.grad.transpose(0, 1)) if layout == "rtd" else (tu.grad, ti.grad)
    assert_parity(du[k], ref_du, "d user_vector[k]")
    assert_parity(di[k], ref_di, "d item_vector[k]")
    for j in range(T):                                   # the other intervals get no gradient
        if j != k:
            assert not du[j].any() and not di[j].any()
    # the default backward is the atomic-free one: bit-identical when repeated; the atomic variant agrees to rounding
    first = (tu.grad.clone(), ti.grad.clone())
    for det in (True, False):
        tu.grad = None; ti.grad = None
        sg.pair_scores(tu, ti, k, torch.from_numpy(uids).cuda(), torch.from_numpy(iids).cuda(), activation=act,
                       leaky=0.5, layout=layout, deterministic=det).backward(torch.from_numpy(g).cuda())
        if det:
            assert torch.equal(tu.grad, first[0]) and torch.equal(ti.grad, first[1])
        else:
            du2 = tu.grad.transpose(0, 1) if layout == "rtd" else tu.grad
            assert_parity(du2[k], ref_du, "d user_vector[k] (atomics)")
    with pytest.raises(IndexError):
        sg.pair_scores(tu, ti, k, torch.tensor([U], device="cuda"), torch.tensor([0], device="cuda"), layout=layout)
    empty = torch.empty(0, dtype=torch.int32, device="cuda")
    assert sg.pair_scores(tu, ti, k, empty, empty, layout=layout).numel() == 0


DOWNSTREAM = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "downstream_*.npz")))


@pytest.mark.parametrize("path", DOWNSTREAM, ids=[os.path.basename(p)[11:-4] for p in DOWNSTREAM])
def test_pair_scores_match_reference_model_py_executed(path):
    """N2 pinned to the reference text: ``preds_one`` of every interval (model.py:197-199) and the plain ``preds``
    (model.py:170-172) as the reference's own lines compute them (tests/golden/make_golden_downstream.py executes
    model.py:111-112,133-203 over the numpy TF stand-in) vs ``sagnn_pair_scores_fwd`` on both layouts; then
    the consumer chain on the GPU -- interval fusion, meta weights, the gathered scores, the hinge -- against the
    executed ``sslloss``."""
    from sagnn_b200.fusion import IntervalFusion, SslHead
    fx = np.load(path)
    T, d, heads, leaky = int(fx["T"]), int(fx["d"]), int(fx["heads"]), float(fx["leaky"])
    uv, iv = torch.from_numpy(fx["user_vector"]).cuda(), torch.from_numpy(fx["item_vector"]).cuda()
    ids = [(torch.from_numpy(fx["suids%d" % k]).cuda(), torch.from_numpy(fx["siids%d" % k]).cuda()) for k in range(T)]
    for layout in ("trd", "rtd"):
        tu, ti = (uv, iv) if layout == "trd" else (uv.transpose(0, 1).contiguous(), iv.transpose(0, 1).contiguous())
        for k in range(T):
            s = sg.pair_scores(tu, ti, k, ids[k][0], ids[k][1], activation="leakyRelu", leaky=leaky, layout=layout)
            assert_parity(s, fx["preds_one%d" % k], "preds_one[%d] (%s)" % (k, layout))
    fu = torch.from_numpy(fx["final_user_vector"].astype(np.float32)).cuda()[None].contiguous()
    fi = torch.from_numpy(fx["final_item_vector"].astype(np.float32)).cuda()[None].contiguous()
    s = sg.pair_scores(fu, fi, 0, torch.from_numpy(fx["uids"]).cuda(), torch.from_numpy(fx["iids"]).cuda(), activation=None)
    assert_parity(s, fx["preds_dot"], "preds (model.py:169-172)")
    # the whole consumer chain in fp32 on the GPU
    m = IntervalFusion(d, heads=heads, device="cuda")
    head = SslHead(d, ssldim=int(fx["ssldim"]), leaky=leaky, device="cuda")
    v = lambda j: torch.from_numpy(fx["var%02d" % j])
    order = ("ln_beta", "ln_gamma", "wq", "bq", "wk", "bk", "wv", "bv")
    # creation order of the executed text: posEmbed, LSTM (1, 2), user block (3..10), item block (11..18), the sequence
    # branch (2 + 2 + 8 per attention layer), then meta2, meta2Bias, meta3, meta3Bias (tests/test_fusion.py::_layout)
    meta0 = 23 + 8 * int(fx["att_layer"])
    assert str(fx["var_names"][meta0]).startswith("meta2") and str(fx["var_names"][1]).startswith("basic_lstm_cell/kernel")
    with torch.no_grad():
        m.lstm_kernel.copy_(v(1)); m.lstm_bias.copy_(v(2))
        for side, o in (("user", 3), ("item", 11)):
            sp_ = m.side_params(side)
            for j, name in enumerate(order):
                sp_[name].copy_(v(o + j))
        for j, p in enumerate((head.meta2, head.meta2_bias, head.meta3, head.meta3_bias)):
            p.copy_(v(meta0 + j).reshape(p.shape))
    gu, gi = m(uv.transpose(0, 1).contiguous(), iv.transpose(0, 1).contiguous())
    scale = float(np.abs(fx["final_user_vector"]).max())
    assert float((gu.detach().cpu().double() - torch.from_numpy(fx["final_user_vector"])).abs().max()) <= 5e-5 * scale
    assert float((gi.detach().cpu().double() - torch.from_numpy(fx["final_item_vector"])).abs().max()) <= 5e-5 * scale
    w = head.user_weight(gu, uv)
    loss = 0.0
    for k in range(T):
        su, si = ids[k]
        final_scores = sg.pair_scores(gu[None].contiguous(), gi[None].contiguous(), 0, su, si, activation="leakyRelu", leaky=leaky)
        interval_scores = sg.pair_scores(uv, iv, k, su, si, activation="leakyRelu", leaky=leaky)
        loss = loss + head.hinge(w[k][su.long()], final_scores, interval_scores)
    assert abs(float(loss.detach()) - float(fx["sslloss"])) <= 1e-4 * abs(float(fx["sslloss"]))


def test_pair_scores_feed_the_propagation_backward():
    """End of the chain the reference builds (model.py:118-129 -> 194-198): the SSL scores of every
    interval on top of `propagate`, gradients w.r.t. the embedding tables through both hand-written
    backward kernels, vs the oracle chain."""
    mats = random_interval_mats(2, 60, 45, 400, seed=17)
    adj, tp = adj_lists(mats)
    T, U, I, d, L, n = 2, 60, 45, 64, 2, 500
    uE, iE, _, _ = random_tables(T, U, I, d, seed=19, scale=0.3)
    rng = np.random.default_rng(23)
    ids = [(rng.integers(0, U, n).astype(np.int32), rng.integers(0, I, n).astype(np.int32)) for _ in range(T)]
    gs = [rng.standard_normal(n).astype(np.float32) for _ in range(T)]
    plan = sg.build_plan(mats)
    u = torch.from_numpy(uE).cuda().requires_grad_(True)
    i = torch.from_numpy(iE).cuda().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, 0.5)
    scores = [sg.pair_scores(uv, iv, k, torch.from_numpy(ids[k][0]).cuda(), torch.from_numpy(ids[k][1]).cuda())
              for k in range(T)]
    torch.autograd.backward(scores, [torch.from_numpy(x).cuda() for x in gs])
    torch.cuda.synchronize()
    ruv, riv, tape = po.propagate_forward(adj, tp, uE.astype(np.float64), iE.astype(np.float64), L, 0.5)
    gU, gI = np.zeros_like(ruv), np.zeros_like(riv)
    for k in range(T):
        assert_parity(scores[k], po.pair_scores(ruv[k], riv[k], ids[k][0], ids[k][1], 0.5)[0], "scores[%d]" % k)
        gU[k], gI[k] = po.pair_scores_backward(ruv[k], riv[k], ids[k][0], ids[k][1], gs[k].astype(np.float64), 0.5)
    rdu, rdi = po.propagate_backward(adj, tp, tape, gU, gI, L, 0.5)
    assert_parity(u.grad, rdu, "dU through pair scores")
    assert_parity(i.grad, rdi, "dI through pair scores")


# ---------------------------------------------------------------- row sharding (second way)
@pytest.mark.parametrize("W,L,d,U,I", [(2, 2, 64, 300, 40), (3, 3, 128, 301, 41), (4, 1, 64, 120, 90)])
def test_row_sharded_virtual_ranks_equal_single_plan_bitwise(W, L, d, U, I):
    """W row-block plans on ONE GPU play the ranks of a row-sharded job: the real stage lists of
    sagnn_b200.dist drive the real C-ABI calls (sagnn_propagate_fwd_layers / _bwd_levels), the table
    all-gather is emulated by copying row blocks between the ranks' workspaces.  Every rank writes
    only its own rows (the others stay NaN until exchanged) and the result is bitwise the
    single-plan one, slices of long rows included."""
    from sagnn_b200 import dist as sd
    T = 2
    mats = random_interval_mats(T, U, I, min(6000, U * I // 3), seed=8)     # item rows of ~150 edges: slices
    ref_plan = sg.build_plan(mats)
    uE, iE, gU, gI = random_tables(T, U, I, d, seed=13)
    ref = run_gpu(ref_plan, uE, iE, gU, gI, L)
    Up, Ip = sd.padded_rows(U, W), sd.padded_rows(I, W)
    bu, bi = Up // W, Ip // W

    def pad(x, rows):
        out = torch.zeros((T, rows, d), dtype=torch.float32, device="cuda")
        out[:, :x.shape[1]] = torch.from_numpy(x).cuda()
        return out

    def exchange(tables, b):                   # tables[r]: rank r's copy; rows [r*b, (r+1)*b) are its own
        for r in range(W):
            for q in range(W):
                if q != r:
                    tables[q][:, r * b:(r + 1) * b] = tables[r][:, r * b:(r + 1) * b]

    nan = lambda rows: torch.full((T, rows, d), float("nan"), dtype=torch.float32, device="cuda")
    bes, outs, grads = [], [], []
    for r in range(W):
        plan = sg.build_plan(mats, U, I, latdim=d, padded_shape=(Up, Ip),
                             row_block=(r * bu, (r + 1) * bu, r * bi, (r + 1) * bi))
        be = sd._CudaRowBackend(plan, L, d, 0.5)
        outs.append((nan(Up), nan(Ip)))
        grads.append((nan(Up), nan(Ip)))
        be.begin_forward(pad(uE, Up), pad(iE, Ip), outs[r][0], outs[r][1], True)
        be.begin_backward(pad(gU, Up), pad(gI, Ip), grads[r][0], grads[r][1])
        bes.append(be)
    for l, ex in sd.forward_stages(L):
        for be in bes:
            be.fwd_layers(l, l + 1)
        if ex is not None:
            exchange([be.table(0, ex)[0] for be in bes], bu)
            exchange([be.table(0, ex)[1] for be in bes], bi)
    torch.cuda.synchronize()
    for r in range(W):                          # before the hand-off only the owned rows exist
        own = torch.zeros(Up, dtype=torch.bool, device="cuda")
        own[r * bu:(r + 1) * bu] = True
        assert torch.isnan(outs[r][0][:, ~own]).all() and not torch.isnan(outs[r][0][:, own]).any()
    exchange([o[0] for o in outs], bu)
    exchange([o[1] for o in outs], bi)
    for ph, ex in sd.backward_stages(L):
        for be in bes:
            be.bwd_levels(ph, ph + 1)
        if ex is not None:
            exchange([be.table(1, ex)[0] for be in bes], bu)
            exchange([be.table(1, ex)[1] for be in bes], bi)
    exchange([g[0] for g in grads], bu)
    exchange([g[1] for g in grads], bi)
    torch.cuda.synchronize()
    for r in range(W):
        assert torch.equal(outs[r][0][:, :U], ref[0]) and torch.equal(outs[r][1][:, :I], ref[1])
        assert torch.equal(grads[r][0][:, :U], ref[2]) and torch.equal(grads[r][1][:, :I], ref[3])
    st = bes[0].plan.stats()
    assert st["short_rows"] + st["long_rows"] == T * (bu + bi)             # only the owned rows are scheduled


def test_row_block_argument_checks():
    mats = random_interval_mats(1, 30, 20, 100, seed=2)
    with pytest.raises(sg.SagnnError):
        sg.build_plan(mats, row_block=(10, 5, 0, 20))
    with pytest.raises(sg.SagnnError):
        sg.build_plan(mats, row_block=(0, 31, 0, 20))
    with pytest.raises(ValueError):
        sg.build_plan(mats, padded_shape=(29, 20))
    plan = sg.build_plan(mats, row_block=(0, 30, 0, 20))                  # the whole graph as one block
    uE, iE, gU, gI = random_tables(1, 30, 20, 64, seed=1)
    a = run_gpu(plan, uE, iE, gU, gI, 2)
    b = run_gpu(sg.build_plan(mats), uE, iE, gU, gI, 2)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_row_sharded_two_gpus_nccl():
    """scripts/row_shard_check.py under torchrun: RowShardedPropagation over NCCL == single-GPU propagate."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(root, "scripts", "row_shard_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ROW_SHARD_OK" in r.stdout


# ---------------------------------------------------------------- BASELINE shape families
@pytest.mark.parametrize("name,scale", [("gowalla", 0.05), ("amazon-book", 0.05), ("amazon-ref", 0.25), ("ml10m", 0.03)])
def test_baseline_shapes_reduced_scale(name, scale):
    g = dh.make_named(name, seed=100, scale=scale)
    check_against_oracle(g.sub_mat, g.meta["d"], g.meta["L"], seed=100, scale=0.05)


def test_config1_trn_mat_time_plumbing(tmp_path):
    """BASELINE config 1 (plumbing): a `trn_mat_time` pickle in the reference's on-disk layout
    (preprocess_to_trnmat.ipynb:1896) -> LoadData mirror -> transToLsts lists -> plan -> propagate
    with the gowalla.sh settings (graphNum 3, gnn_layer 2, latdim 64, leaky 0.5), at reduced scale."""
    g = dh.make_named("gowalla", seed=100, scale=0.04)
    path = str(tmp_path / "trn_mat_time")
    dh.write_trn_mat_time(path, g)
    h = dh.load_trn_mat_time(path, graph_num=3)                 # args.user, args.item from trnMat[0].shape
    assert (h.n_user, h.n_item) == (g.n_user, g.n_item)
    lists = [sg.transToLsts(m, norm=True) for m in h.sub_mat]   # what model.py:230-233 feeds TF
    assert all(np.all(d == 0) for _, d, _ in lists)             # the dead normalisation
    plan = sg.build_plan([l[0] for l in lists], U=h.n_user, I=h.n_item)
    check_against_oracle(h.sub_mat, 64, 2, leaky=0.5, seed=100, scale=0.05)
    # the plan built from adjacency lists equals the one built from the matrices
    plan2 = sg.build_plan(h.sub_mat)
    # ... and so does the plan built from the memory-mapped binary container
    bpath = str(tmp_path / "trn_mat_time.sagnncsr")
    dh.write_trn_mat_bin(bpath, g)
    plan3 = sg.build_plan(dh.load_trn_mat_bin(bpath, graph_num=3).sub_mat)
    for k in range(3):
        for side in (0, 1):
            assert torch.equal(plan.adjacency_list(k, side), plan2.adjacency_list(k, side))
            assert torch.equal(plan.adjacency_list(k, side), plan3.adjacency_list(k, side))
            assert torch.equal(plan2.degrees(k, side, value_sum=True)[1], plan3.degrees(k, side, value_sum=True)[1])


@pytest.mark.parametrize("name", ["amazon-book", "ml10m"])
def test_other_baseline_shapes_full_size(name):
    """BASELINE configs 3 and 4 at full size (three-part parity vs the fused C oracle, fp64)."""
    g = dh.make_named(name, seed=100)
    L, d = g.meta["L"], g.meta["d"]
    T, U, I = g.graph_num, g.n_user, g.n_item
    uE = dh.xavier_embeddings(T, U, d, 100)
    iE = dh.xavier_embeddings(T, I, d, 101)
    rng = np.random.default_rng(100)
    gU = rng.standard_normal((T, U, d), dtype=np.float32)
    gI = rng.standard_normal((T, I, d), dtype=np.float32)
    check_against_oracle(g.sub_mat, d, L, tables=(uE, iE, gU, gI), max_ties=256, report="%s full size (L=%d, d=%d)" % (name, L, d))


def test_many_tiny_rows_batched_packet_queue():
    """Hundreds of thousands of 1-2 edge rows (the scaled-s10 regime): a warp has hundreds of packets, so the
    packet queue hands them out in batches of several per atomic; packets of partly empty rows, long tails of
    zero-degree rows, the last partial packet of every segment."""
    rng = np.random.default_rng(77)
    U, I, E = 300_000, 120_000, 380_000
    keys = np.unique(rng.integers(0, U * I, size=E, dtype=np.int64))
    m = sp.csr_matrix((np.ones(keys.size, np.intc), (keys // I, keys % I)), shape=(U, I))
    check_against_oracle([m], 64, 2, seed=5, scale=0.5, max_ties=64)


def test_monster_rows_three_level_slice_tree():
    """Rows far beyond 16*16*64 edges (as in BASELINE config 5, where head items collect millions
    of users) are reduced through a three-level ticket tree: a hub item linked to 90 % of 120 K
    users and a hub user linked to every item, on top of a sparse random graph."""
    U, I = 120_000, 3000
    rng = np.random.default_rng(5)
    n = 600_000
    r = rng.integers(0, U, n)
    c = rng.integers(0, I, n)
    hub_users = np.flatnonzero(rng.random(U) < 0.9)
    r = np.concatenate([r, hub_users, np.full(I, 77)])
    c = np.concatenate([c, np.full(hub_users.size, 5), np.arange(I)])
    m = sp.csr_matrix((np.ones(r.size, np.intc), (r, c)), shape=(U, I))
    m.data[:] = 1
    m.sort_indices()
    plan = check_against_oracle([m], 64, 2, seed=9, scale=0.02, max_ties=64)
    st = plan.stats()
    assert st["max_degree"] > 16 * 16 * 64 and st["chunks"] > 1700


def test_gowalla_full_size_parity_and_properties():
    """BASELINE config 2 at full size: three-part parity vs the (fast, fused) C oracle in fp64,
    plus size-independent properties: the backward is linear in the upstream gradient and
    edgeless rows return (L+1) * embedding."""
    g = dh.make_named("gowalla", seed=100)
    T, U, I, d, L = 3, g.n_user, g.n_item, 64, 2
    uE = dh.xavier_embeddings(T, U, d, 100)
    iE = dh.xavier_embeddings(T, I, d, 101)
    rng = np.random.default_rng(100)
    gU = rng.standard_normal((T, U, d)).astype(np.float32)
    gI = rng.standard_normal((T, I, d)).astype(np.float32)
    plan = check_against_oracle(g.sub_mat, d, L, tables=(uE, iE, gU, gI), max_ties=64, report="gowalla full size (L=%d, d=%d)" % (L, d))
    got = run_gpu(plan, uE, iE, gU, gI, L)
    deg = plan.degrees(0, 0).cpu().numpy()
    lonely = np.flatnonzero(deg == 0)[:50]
    if lonely.size:
        np.testing.assert_allclose(got[0][0].detach().cpu().numpy()[lonely], 3 * uE[0][lonely], rtol=1e-6)
    # backward is linear in the upstream for a fixed forward: bwd(2g) == 2 bwd(g) exactly (power of 2)
    got2 = run_gpu(plan, uE, iE, 2 * gU, 2 * gI, L)
    assert torch.equal(got2[2], 2 * got[2]) and torch.equal(got2[3], 2 * got[3])
