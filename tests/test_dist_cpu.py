"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: interval assignment, the ordered
all-gather of per-interval outputs and its backward.  The local compute is stood in for by the
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
        out_q.put((rank, "ok" if torch.equal(got, want) else "mismatch", ""))
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
