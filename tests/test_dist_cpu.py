"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: interval assignment, the ordered
all-gather of per-interval outputs and its backward, and the row-sharded schedule (per-layer
table all-gathers).  The local compute is stood in for by the
oracle here (tests may use it); on GPUs it is the CUDA path (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sagnn_b200 import dist as sd


def test_assign_intervals_lpt():
    assert sd.assign_intervals([10, 10, 6], 1) == [0, 0, 0]
    own = sd.assign_intervals([10, 10, 6], 2)
    assert sorted(own) == [0, 0, 1] or sorted(own) == [0, 1, 1]
    loads = [sum(n for n, o in zip([10, 10, 6], own) if o == r) for r in range(2)]
    assert max(loads) == 16
    own = sd.assign_intervals([5, 4, 3, 2, 1], 8)
    assert len(set(own)) == 5                                  # T <= world: one interval per rank
    assert sd.assign_intervals([7, 7, 7, 7], 2) == sd.assign_intervals([7, 7, 7, 7], 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from oracle import propagate_oracle as po
        from helpers import adj_lists, random_interval_mats, random_tables
        U, I, d, L = 40, 30, 8, 2
        mats = random_interval_mats(T, U, I, 150, seed=21)
        # uneven interval sizes so that the LPT assignment is not trivial
        mats[0] = random_interval_mats(1, U, I, 400, seed=22)[0]
        adj, tp = adj_lists(mats)
        uE, iE, gU, gI = [x.astype(np.float64) for x in random_tables(T, U, I, d, seed=5)]
        owners = sd.assign_intervals([m.nnz for m in mats], world)
        mine = sd.local_intervals(owners, rank)
        # local compute (oracle stands in for the CUDA path on CPU)
        uv_l, iv_l, tape = po.propagate_forward([adj[k] for k in mine], [tp[k] for k in mine],
                                                uE[mine], iE[mine], L, 0.5)
        uv_t = torch.from_numpy(uv_l).requires_grad_(True)
        iv_t = torch.from_numpy(iv_l).requires_grad_(True)
        full_u = sd.gather_intervals(uv_t, owners, rank)
        full_i = sd.gather_intervals(iv_t, owners, rank)
        ref_u, ref_i, _ = po.propagate_forward(adj, tp, uE, iE, L, 0.5)
        ok_fwd = np.array_equal(full_u.detach().numpy(), ref_u) and np.array_equal(full_i.detach().numpy(), ref_i)
        # replicated consumer: upstream gradient of the gathered tensor -> local slice
        torch.autograd.backward([full_u, full_i], [torch.from_numpy(gU), torch.from_numpy(gI)])
        ok_bwd = np.array_equal(uv_t.grad.numpy(), gU[mine]) and np.array_equal(iv_t.grad.numpy(), gI[mine])
        du_l, di_l = po.propagate_backward([adj[k] for k in mine], [tp[k] for k in mine], tape,
                                           uv_t.grad.numpy(), iv_t.grad.numpy(), L, 0.5)
        du_ref, di_ref = po.propagate_backward(adj, tp, po.propagate_forward(adj, tp, uE, iE, L, 0.5)[2], gU, gI, L, 0.5)
        ok_grad = np.array_equal(du_l, du_ref[mine]) and np.array_equal(di_l, di_ref[mine])
        out_q.put((rank, ok_fwd, ok_bwd, ok_grad, mine))
    finally:
        dist.destroy_process_group()


def _a2a_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, R, d = 2, 7, 3                                   # R not divisible by world: padding path
        full = torch.arange(world * T * R * d, dtype=torch.float32).view(world, T, R, d)
        try:
            got = sd.exchange_rows(full[rank].clone())
        except Exception as e:                              # gloo builds without all_to_all
            out_q.put((rank, "skip", str(e)))
            return
        size, blocks = sd.row_blocks(R, world)
        lo, hi = blocks[rank]
        want = torch.zeros(world * T, size, d)
        want[:, :hi - lo] = full[:, :, lo:hi].reshape(world * T, hi - lo, d)
        ok = torch.equal(got, want)
        # zero-copy variant: the [R_pad, T, d] (rtd) output buffer is the send buffer
        Rp = size * world
        full_rtd = torch.zeros(world, Rp, T, d)
        full_rtd[:, :R] = full.transpose(1, 2)
        got2 = sd.exchange_rows_rtd(full_rtd[rank].clone())                     # [src, block, T, d]
        want2 = full_rtd[:, rank * size:(rank + 1) * size]
        ok = ok and torch.equal(got2, want2) and torch.equal(got2.transpose(1, 2).reshape(world * T, size, d), want)
        out_q.put((rank, "ok" if ok else "mismatch", ""))
    finally:
        dist.destroy_process_group()


def test_row_block_exchange_gloo_world2():
    """The all-to-all hand-off to a row-sharded consumer delivers, on every rank, its row block of
    every rank's intervals (rank order), with zero padding."""
    assert sd.row_blocks(7, 2) == (4, [(0, 4), (4, 7)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    if any(r[1] == "skip" for r in res):
        pytest.skip("gloo without all_to_all: " + res[0][2])
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.parametrize("T", [3, 5, 1])
def test_interval_sharding_gloo_world2(T):
    """1-vs-2 rank equality of outputs and gradients is bitwise (interval sharding)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, ok_fwd, ok_bwd, ok_grad, mine in res:
        assert ok_fwd and ok_bwd and ok_grad, (rank, ok_fwd, ok_bwd, ok_grad)
        covered += mine
    assert sorted(covered) == list(range(T))


def _row_worker(rank, world, port, L, U, I, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from oracle import propagate_oracle as po
        from helpers import DenseRowBackend, adj_lists, random_interval_mats, random_tables
        T, d = 2, 8
        mats = random_interval_mats(T, U, I, 260, seed=33)
        adj, tp = adj_lists(mats)
        uE, iE, gU, gI = [x.astype(np.float64) for x in random_tables(T, U, I, d, seed=6)]
        class _CpuRowSharded(sd.RowShardedPropagation):      # test-only stand-in for the two CUDA hooks
            def _build_plan(self, sub_mats, device, latdim):
                return None

            def _make_backend(self, d_):
                return DenseRowBackend(self, d_, mats)

        rs = _CpuRowSharded(mats, U, I, n_layers=L, leaky=0.5)
        u = torch.from_numpy(uE).requires_grad_(True)
        i = torch.from_numpy(iE).requires_grad_(True)
        uv, iv = rs(u, i)
        torch.autograd.backward([uv, iv], [torch.from_numpy(gU), torch.from_numpy(gI)])
        ref = po.propagate(adj, tp, uE, iE, gU, gI, L, 0.5, np.float64)
        got = [uv.detach().numpy(), iv.detach().numpy(), u.grad.numpy(), i.grad.numpy()]
        err = [float(np.max(np.abs(g - r)) / np.max(np.abs(r))) for g, r in zip(got, ref)]
        finite = all(np.isfinite(g).all() for g in got)
        out_q.put((rank, rs.row_block, (rs.U_pad, rs.I_pad), err, finite))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,U,I", [(1, 40, 30), (2, 41, 29), (3, 40, 31)])
def test_row_sharding_schedule_gloo_world2(L, U, I):
    """The row-sharded schedule of sagnn_b200.dist (stage lists, in-place table all-gathers, padding
    to equal row blocks, autograd wiring) reproduces the unsharded oracle on 2 ranks; the C-ABI calls
    are stood in for by a dense CPU backend that, like the kernels, writes owned rows only (every
    other row starts as NaN, so a missing exchange cannot go unnoticed)."""
    assert sd.forward_stages(3) == [(0, 0), (1, 1), (2, None)]
    assert sd.backward_stages(2) == [(0, 0), (1, 1), (2, None)]
    assert sd.padded_rows(41, 2) == 42 and sd.padded_rows(40, 2) == 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, L, U, I, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blocks = sorted(r[1] for r in res)
    assert blocks[0][0] == 0 and blocks[0][1] == blocks[1][0] and blocks[1][1] == res[0][2][0]
    for rank, _, _, err, finite in res:
        assert finite and max(err) < 1e-12, (rank, err)
