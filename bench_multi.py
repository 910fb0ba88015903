"""Multi-GPU measurements on the configurations north_star names (imported by bench.py at N > 1;
also runnable on their own under torchrun).  Nothing here changes the headline line: the results go
into the ``extras`` object of bench.py's JSON line, so that the driver's own 1/2/4/8-GPU runs of
``bench.py --gpus N`` carry them.

  scaled   BASELINE config 5 (10 M users x 2 M items, ~1 B edges over 8 intervals): every rank owns ONE
           125 M-edge interval of that shape (generated on its GPU), i.e. at N = 8 the whole config.
           Intervals share nothing (model.py:108-109,118-129), so the propagation has no collective;
           timed are (a) the local fwd+bwd step, (b) the step with the hand-off fused into the forward
           epilogue (peer stores of row blocks, a row-sharded consumer), (c) the step with the NCCL
           all-gather of the layer outputs to a replicated consumer.  Scaling = N * t_one_interval_one_GPU
           / t_step: the throughput of N GPUs over one GPU working through the N intervals one by one.
  amazon   BASELINE config 3 (Amazon-book shape, T = 5), STRONG scaling: rank 0 first times the whole
           workload alone, then the intervals are dealt over the ranks (LPT) and every rank runs its
           share with the all-gather of the outputs; the gathered result is compared bitwise with the
           single-GPU one.  T = 5 intervals bound the speed-up by 5 (ranks beyond that stay idle).
  rowshard the row-sharded propagation (per-layer table all-gathers) on a small graph, bitwise against the
           single-GPU result -- the check the 1-GPU test box has to skip.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch


def _timed(fn, steps, warmup, dist, dev, pre=None):
    """CUDA-event ms per call, max over ranks; ``pre`` runs before every call outside the event pair."""
    for _ in range(warmup):
        if pre:
            pre()
        fn()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if pre:
            pre()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _alg_bytes(E, U, I, d, L):
    """B_alg (perfect cache) and B_nr (no gather reuse) of one interval, fwd+bwd."""
    n = U + I
    b_alg = L * (16 * E + 8 * (n + 2) + 36.125 * d * n)
    b_nr = L * (4 * E * (4 * d + 4) + 8 * (n + 2) + 28.125 * d * n)
    return b_alg, b_nr


def scaled_config(args, rank, world, dev, dist, peak_gbs):
    import sagnn_b200 as sg
    from sagnn_b200 import data_handler as dh
    from sagnn_b200.step import PropagationStep
    s = dict(dh.SHAPES["scaled"])
    sc = float(args.extras_scale)
    U, I, L, d = max(64, int(s["U"] * sc)), max(64, int(s["I"] * sc)), s["L"], s["d"]
    e_target = int(s["E"] / s["T"] * sc)
    t0 = time.perf_counter()
    row, col = dh.make_interval_device(U, I, e_target, s["au"], s["ai"], seed=args.seed + 17 * rank, device=dev)
    E = int(row.numel())
    torch.cuda.synchronize(dev)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = sg.build_plan([(row, col)], U=U, I=I, device=dev, latdim=d)
    torch.cuda.synchronize(dev)
    plan_s = time.perf_counter() - t0
    del row, col
    step = PropagationStep(plan, L, d, 0.5, layout="rtd", row_multiple=world)
    a_u, a_i = float(np.sqrt(6.0 / (U + d))), float(np.sqrt(6.0 / (I + d)))
    g = torch.Generator(device=dev).manual_seed(args.seed + rank)
    step.u_embed.uniform_(-a_u, a_u, generator=g)
    step.i_embed.uniform_(-a_i, a_i, generator=g)
    step.g_user.normal_(generator=g)
    step.g_item.normal_(generator=g)
    step.calibrate(rounds=1)
    steps, warm = max(3, min(args.steps, 5)), 2
    out = {"shape": {"U": U, "I": I, "intervals_total": world, "intervals_per_rank": 1, "layers": L, "latdim": d,
                     "scale": sc}, "generate_s": gen_s, "plan_build_s": plan_s}
    # (a) compute only: what one GPU needs for one interval
    step.capture()
    t_local = _timed(step.replay, steps, warm, dist, dev)
    e_all = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e_all, op=dist.ReduceOp.SUM)
    e_all = float(e_all.item())
    b_alg, b_nr = _alg_bytes(E, U, I, d, L)
    out["edges_this_rank"] = E
    out["edges_all_ranks"] = e_all
    out["local_step"] = {"ms": t_local, "alg_GBps": b_alg / t_local / 1e6, "alg_frac_of_peak": b_alg / t_local / 1e6 / peak_gbs,
                         "no_reuse_GBps": b_nr / t_local / 1e6, "no_reuse_frac_of_peak": b_nr / t_local / 1e6 / peak_gbs,
                         "note": "fwd+bwd of one interval, CUDA graph, max over ranks; B_alg / B_nr per SURVEY 8(d)"}
    if world == 1 or dist is None:
        return out
    side = torch.cuda.Stream(device=dev)
    # (b) hand-off fused into the forward epilogue (row-sharded consumer): peer stores over NVLink
    try:
        import torch.distributed._symmetric_memory as symm
        bu, bi = step.user_out_full.shape[0] // world, step.item_out_full.shape[0] // world
        rcv_u = symm.empty((world, bu, 1, d), dtype=torch.float32, device=dev)
        rcv_i = symm.empty((world, bi, 1, d), dtype=torch.float32, device=dev)
        hdl_u = symm.rendezvous(rcv_u, dist.group.WORLD)
        hdl_i = symm.rendezvous(rcv_i, dist.group.WORLD)
        step.graph = None
        step.set_scatter(world, rank, list(hdl_u.buffer_ptrs), list(hdl_i.buffer_ptrs))
        step.capture()

        def fused_step():
            step.replay()
            hdl_u.barrier(channel=0)        # every rank's rows have landed in my receive slabs

        t_fused = _timed(fused_step, steps, warm, dist, dev)
        out["fused_handoff_step"] = {"ms": t_fused, "edge_traversals_per_s": 4 * L * e_all / (t_fused * 1e-3),
                                     "scaling_vs_one_gpu": world * t_local / t_fused,
                                     "bytes_sent_per_rank": (world - 1) / world * (U + I) * d * 4}
        step.set_scatter(0, 0, None, None)
        step.graph = None
        del rcv_u, rcv_i
    except Exception as e:   # noqa: BLE001
        out["fused_handoff_step"] = {"error": repr(e)[:300]}
    # (c) NCCL all-gather of the layer outputs to a replicated consumer (north_star's collective)
    try:
        gat_u = torch.empty((world,) + tuple(step.user_out.shape), dtype=torch.float32, device=dev)
        gat_i = torch.empty((world,) + tuple(step.item_out.shape), dtype=torch.float32, device=dev)

        def gather_step():
            step.forward()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                     # overlaps the backward
                dist.all_gather_into_tensor(gat_u, step.user_out)
                dist.all_gather_into_tensor(gat_i, step.item_out)
            step.backward()
            torch.cuda.current_stream().wait_stream(side)

        t_ag = _timed(gather_step, steps, warm, dist, dev)
        recv = (world - 1) * (U + I) * d * 4
        out["allgather_step"] = {"ms": t_ag, "edge_traversals_per_s": 4 * L * e_all / (t_ag * 1e-3),
                                 "scaling_vs_one_gpu": world * t_local / t_ag, "bytes_received_per_rank": recv,
                                 "note": "all-gather issued after the forward on a side stream, overlapping the backward"}
        del gat_u, gat_i
    except Exception as e:   # noqa: BLE001
        out["allgather_step"] = {"error": repr(e)[:300]}
    return out


def amazon_strong(args, rank, world, dev, dist):
    import sagnn_b200 as sg
    from sagnn_b200 import data_handler as dh
    from sagnn_b200 import dist as sd
    from sagnn_b200.step import PropagationStep
    name = "amazon-book"
    g = dh.make_named(name, seed=args.seed, scale=float(args.extras_amazon_scale))   # same graphs on every rank
    T, U, I, L, d = g.graph_num, g.n_user, g.n_item, g.meta["L"], g.meta["d"]
    gen = torch.Generator(device=dev).manual_seed(args.seed)                       # same tables on every rank
    uE = torch.randn((T, U, d), device=dev, generator=gen) * 0.01
    iE = torch.randn((T, I, d), device=dev, generator=gen) * 0.01
    gU = torch.randn((T, U, d), device=dev, generator=gen)
    gI = torch.randn((T, I, d), device=dev, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps, warm = max(5, min(args.steps, 10)), 3
    out = {"shape": {"U": U, "I": I, "T": T, "edges": int(sum(g.nnz)), "layers": L, "latdim": d}}

    def make_step(mats, ks):
        plan = sg.build_plan(mats, device=dev, latdim=d)
        st = PropagationStep(plan, L, d, 0.5)
        idx = torch.tensor(ks, device=dev)
        st.u_embed.copy_(uE.index_select(0, idx)); st.i_embed.copy_(iE.index_select(0, idx))
        st.g_user.copy_(gU.index_select(0, idx)); st.g_item.copy_(gI.index_select(0, idx))
        st.calibrate(rounds=2)
        return st

    # N = 1: the whole workload on this GPU (every rank does it: the max is the honest single-GPU time)
    full = make_step(g.sub_mat, list(range(T)))
    full.capture()
    t1 = _timed(full.replay, steps, warm, dist, dev, pre=flush.zero_)
    ref = [full.user_out.clone(), full.item_out.clone(), full.d_u.clone(), full.d_i.clone()]
    out["one_gpu_ms"] = t1
    if world == 1 or dist is None:
        return out
    assign = sd.assign_intervals(g.nnz, world)          # LPT by edge count
    mine = [k for k in range(T) if assign[k] == rank]
    out["assignment"] = [int(a) for a in assign]
    side = torch.cuda.Stream(device=dev)
    counts = [sum(1 for k in range(T) if assign[k] == r) for r in range(world)]
    maxc = max(counts)
    # every rank sends [maxc, R, d] (its intervals, zero-padded to the largest share): ONE all-gather per side
    send_u = torch.zeros((maxc, U, d), dtype=torch.float32, device=dev)
    send_i = torch.zeros((maxc, I, d), dtype=torch.float32, device=dev)
    gat_u = torch.zeros((world, maxc, U, d), dtype=torch.float32, device=dev)
    gat_i = torch.zeros((world, maxc, I, d), dtype=torch.float32, device=dev)
    loc = make_step([g.sub_mat[k] for k in mine], mine) if mine else None
    if loc is not None:
        loc.user_out = loc.user_out_full = send_u[:len(mine)]     # the epilogue writes straight into the send buffer
        loc.item_out = loc.item_out_full = send_i[:len(mine)]
        loc.capture()

    def sharded_step():
        if loc is not None:
            loc.forward()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                          # hand-off of the outputs overlaps the backward
            dist.all_gather_into_tensor(gat_u, send_u)
            dist.all_gather_into_tensor(gat_i, send_i)
        if loc is not None:
            loc.backward()
        torch.cuda.current_stream().wait_stream(side)

    def compute_only():
        if loc is not None:
            loc.replay()

    t_c = _timed(compute_only, steps, warm, dist, dev, pre=flush.zero_)
    t_n = _timed(sharded_step, steps, warm, dist, dev, pre=flush.zero_)
    # bitwise: slot j of rank r's share == that interval of the single-GPU outputs
    ok = True
    for r in range(world):
        for jslot, k in enumerate([k for k in range(T) if assign[k] == r]):
            ok = ok and torch.equal(gat_u[r, jslot], ref[0][k]) and torch.equal(gat_i[r, jslot], ref[1][k])
    if loc is not None:
        midx = torch.tensor(mine, device=dev)
        ok = ok and torch.equal(loc.d_u, ref[2].index_select(0, midx)) and torch.equal(loc.d_i, ref[3].index_select(0, midx))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out.update({"n_gpu_ms_compute_only": t_c, "n_gpu_ms_with_output_gather": t_n, "speedup_compute_only": t1 / t_c,
                "speedup_with_output_gather": t1 / t_n, "busy_ranks": int(sum(1 for c in counts if c)),
                "gathered_bytes_per_rank": int((world - 1) * maxc * (U + I) * d * 4),
                "bitwise_equal_to_one_gpu": bool(flag.item()),
                "note": "interval sharding (LPT); T=%d intervals bound the speed-up by %d; outputs handed to a replicated "
                        "consumer by one NCCL all-gather per side (shares zero-padded to the largest) on a side stream overlapping the backward" % (T, T)})
    return out


def rowshard_check(args, rank, world, dev, dist):
    import sagnn_b200 as sg
    from sagnn_b200 import data_handler as dh
    from sagnn_b200 import dist as sd
    g = dh.make_named("small", seed=100)
    T, U, I, L, d = g.graph_num, g.n_user, g.n_item, 2, 64
    gen = torch.Generator(device=dev).manual_seed(7)
    mk = lambda rows: torch.randn((T, rows, d), device=dev, generator=gen)
    uE, iE, gU, gI = mk(U).requires_grad_(True), mk(I).requires_grad_(True), mk(U), mk(I)
    rs = sd.RowShardedPropagation(g.sub_mat, U, I, n_layers=L, leaky=0.5, device=dev, latdim=d)
    uv, iv = rs(uE, iE)
    torch.autograd.backward([uv, iv], [gU, gI])
    got = [uv.detach().clone(), iv.detach().clone(), uE.grad.clone(), iE.grad.clone()]
    uE.grad = iE.grad = None
    plan = sg.build_plan(g.sub_mat, device=dev, latdim=d)
    uv1, iv1 = sg.propagate(plan, uE, iE, L, 0.5)
    torch.autograd.backward([uv1, iv1], [gU, gI])
    ok = all(torch.equal(a, b) for a, b in zip(got, [uv1.detach(), iv1.detach(), uE.grad, iE.grad]))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"row_sharded_bitwise_equal": bool(flag.item()), "graph": "small (T=3, U=4000, I=3000, 60 K edges), L=2, d=64",
            "note": "RowShardedPropagation over NCCL (per-layer in-place table all-gathers) vs sagnn_b200.propagate on one GPU"}


def fusion_chain_check(args, rank, world, dev, dist):
    """The data-parallel chain the fused hand-off exists for, over real peer memory: every rank owns two
    interval graphs, runs its forward with sagnn_propagate_fwd_scatter (peer stores into the symmetric-memory
    receive buffers), assembles ITS row block of all intervals from the slabs, runs the interval fusion
    (model.py:135-155) and a loss, sends the dense upstream back to the interval owners (one all-to-all) and
    runs sagnn_propagate_bwd_ex; outputs and gradients are compared with the single-GPU chain."""
    import torch.distributed._symmetric_memory as symm
    import sagnn_b200 as sg
    from sagnn_b200 import data_handler as dh
    from sagnn_b200.fusion import IntervalFusion, slabs_to_rtd
    from sagnn_b200.step import PropagationStep
    tl, U, I, d, L = 2, 3000, 2000, 64, 2
    T = tl * world
    g = dh.make_interval_graphs(U, I, T, [20000] * T, seed=5)                  # same graphs on every rank
    owners = [k // tl for k in range(T)]
    mine = [k for k in range(T) if owners[k] == rank]
    gen = torch.Generator(device=dev).manual_seed(11)                          # same tables on every rank
    uE = torch.randn((T, U, d), device=dev, generator=gen) * 0.3
    iE = torch.randn((T, I, d), device=dev, generator=gen) * 0.3
    wu = torch.randn((U, d), device=dev, generator=gen)
    wi = torch.randn((I, d), device=dev, generator=gen)
    fusion = IntervalFusion(d, heads=16, device=dev, seed=5)
    # single-GPU chain
    plan = sg.build_plan(g.sub_mat, device=dev, latdim=d)
    u, i = uE.clone().requires_grad_(True), iE.clone().requires_grad_(True)
    uv, iv = sg.propagate(plan, u, i, L, 0.5, layout="rtd")
    fu, fi = fusion(uv, iv)
    ((fu * wu).sum() + (fi * wi).sum()).backward()
    ref = [fu.detach(), fi.detach(), u.grad, i.grad]
    fusion.zero_grad()
    # sharded chain
    st = PropagationStep(sg.build_plan([g.sub_mat[k] for k in mine], device=dev, latdim=d), L, d, layout="rtd",
                         row_multiple=world)
    st.u_embed.copy_(uE[mine]); st.i_embed.copy_(iE[mine])
    bu, bi = st.user_out_full.shape[0] // world, st.item_out_full.shape[0] // world
    rcv_u = symm.empty((world, bu, tl, d), dtype=torch.float32, device=dev)
    rcv_i = symm.empty((world, bi, tl, d), dtype=torch.float32, device=dev)
    rcv_u.zero_(); rcv_i.zero_()
    hu, hi = symm.rendezvous(rcv_u, dist.group.WORLD), symm.rendezvous(rcv_i, dist.group.WORLD)
    hu.barrier(channel=0)
    st.set_scatter(world, rank, list(hu.buffer_ptrs), list(hi.buffer_ptrs))
    st.forward()
    hu.barrier(channel=0)
    torch.cuda.synchronize(dev)
    lo_u, hi_u, lo_i, hi_i = rank * bu, min((rank + 1) * bu, U), rank * bi, min((rank + 1) * bi, I)
    xu = slabs_to_rtd(rcv_u, owners, max(hi_u - lo_u, 0)).clone().requires_grad_(True)
    xi = slabs_to_rtd(rcv_i, owners, max(hi_i - lo_i, 0)).clone().requires_grad_(True)
    fu, fi = fusion(xu, xi)
    relerr = lambda a, b: float(((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30)).item())
    e_f = max(relerr(fu, ref[0][lo_u:hi_u]), relerr(fi, ref[1][lo_i:hi_i]))
    ((fu * wu[lo_u:hi_u]).sum() + (fi * wi[lo_i:hi_i]).sum()).backward()
    # reverse hand-off: my block's upstream of rank r's intervals -> rank r
    send_u = torch.zeros((world, bu, tl, d), device=dev); send_i = torch.zeros((world, bi, tl, d), device=dev)
    for r in range(world):
        ks = [k for k in range(T) if owners[k] == r]
        send_u[r, :hi_u - lo_u] = xu.grad[:, ks]
        send_i[r, :hi_i - lo_i] = xi.grad[:, ks]
    back_u, back_i = torch.empty_like(send_u), torch.empty_like(send_i)
    dist.all_to_all_single(back_u.view(world, -1), send_u.view(world, -1))
    dist.all_to_all_single(back_i.view(world, -1), send_i.view(world, -1))
    st.set_scatter(0, 0, None, None)
    st.g_user.copy_(back_u.reshape(world * bu, tl, d)[:U]); st.g_item.copy_(back_i.reshape(world * bi, tl, d)[:I])
    st.backward()
    torch.cuda.synchronize(dev)
    e_g = max(relerr(st.d_u, ref[2][mine]), relerr(st.d_i, ref[3][mine]))
    t = torch.tensor([e_f, e_g], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_f, e_g = float(t[0]), float(t[1])
    return {"chain_matches_single_gpu": bool(e_f <= 1e-5 and e_g <= 1e-5), "max_relerr_fused_vectors": e_f,
            "max_relerr_embedding_grads": e_g,
            "graph": "T=%d intervals of 20 K edges (2 per rank), U=%d, I=%d, L=%d, d=%d" % (T, U, I, L, d),
            "note": "propagate (fwd_scatter, peer stores) -> slabs -> LSTM/LN/MHSA/mean per row block -> loss -> all-to-all "
                    "of the [blk,T,d] upstream -> sagnn_propagate_bwd_ex; vs propagate(layout=rtd) -> fusion -> autograd on one GPU"}


def run_extras(args, rank, world, dev, dist, peak_gbs):
    """Every part is independent and failure-tolerant: an exception becomes an ``error`` entry."""
    res = {}
    parts = [("rowshard", lambda: rowshard_check(args, rank, world, dev, dist)),
             ("fusion_chain", lambda: fusion_chain_check(args, rank, world, dev, dist)),
             ("amazon_book_strong_scaling", lambda: amazon_strong(args, rank, world, dev, dist)),
             ("scaled_config5", lambda: scaled_config(args, rank, world, dev, dist, peak_gbs))]
    for name, fn in parts:
        if world == 1 and name in ("rowshard", "fusion_chain"):
            continue
        t0 = time.perf_counter()
        try:
            res[name] = fn()
        except Exception as e:   # noqa: BLE001
            res[name] = {"error": repr(e)[:400]}
        res[name]["wall_s"] = time.perf_counter() - t0
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()
        if dist is not None:
            dist.barrier()
    return res
