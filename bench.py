#!/usr/bin/env python
"""bench.py -- fwd+bwd interval-graph propagation throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gowalla] [--impl ours|reference]

A "step" is one forward + one backward of the propagation over all T interval graphs and L
layers of the workload: 4*L*sum(E_k) edge traversals.  Own arm (default):
  value      edge traversals/s with all inputs resident in HBM (CUDA events, L2 flushed between
             steps, whole step replayed from one CUDA graph);
  e2e        the same metric through the C-ABI host entry point (sagnn_propagate_host) with
             pinned HOST buffers: H2D of embeddings + upstream grads and D2H of outputs + grads
             inside the timed region;
  roofline   per-launch algorithmic bytes / measured duration of the layer kernel;
  cpu_baseline  the TF1-graph restatement (oracle/tf1_mirror.py, torch CPU, all host threads).
Reference arm (--impl reference): times that CPU restatement only (the reference itself needs
TensorFlow 1.14, which cannot be installed here).
Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank owns one full set of T
interval graphs (interval sharding: intervals share nothing), no data-path collective in the
propagation itself.  The hand-off of the per-rank outputs runs on a side stream overlapped with
the backward: by default one NCCL all-to-all of row blocks to a row-sharded consumer, sent straight
from the epilogue's [R,T,d] output (no pack copies); --exchange allgather = the all-gather to a
replicated consumer that north_star names; --exchange none = compute only.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gowalla")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--latdim", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-allgather", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "ce", "allgather", "alltoall", "none"],
                    help="multi-GPU hand-off of the outputs: all-to-all of row blocks (row-sharded consumer; the "
                         "default for N > 1, sent straight from the [R,T,d] epilogue output), all-gather "
                         "(replicated consumer, the collective north_star names), fused (the same row-block "
                         "hand-off written by the epilogue itself into the peers' symmetric-memory receive buffers: "
                         "peer stores over NVLink, no collective kernel), ce (the same receive buffers filled by the "
                         "copy engines: one peer memcpy per row block on a side stream while the SMs run the backward) or none")
    ap.add_argument("--layout", default=None, choices=["trd", "rtd"],
                    help="rtd: outputs / upstream gradients in the [R,T,d] layout of model.py:133-134 (fused "
                         "transpose); default trd, rtd with --exchange alltoall")
    ap.add_argument("--no-calibrate", action="store_true", help="keep the static cost-model CTA split")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic only: keep L2 warm between steps")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--extras", default="auto", choices=["auto", "on", "off"],
                    help="N > 1 (auto) / on: also measure BASELINE configs 3 and 5 the multi-GPU way north_star names "
                         "(bench_multi.py) and run the row-sharding bitwise check; results go into line['extras']")
    ap.add_argument("--extras-scale", type=float, default=1.0, help="size factor of the config-5 interval (tests)")
    ap.add_argument("--extras-amazon-scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=100)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(nnz, U, I, d, L):
    """algorithmic (compulsory, perfect-cache) bytes."""
    n = U + I
    fwd_layer = sum(8 * e + 4 * (n + 2) + 20 * d * n for e in nnz)
    bwd_layer = sum(8 * e + 4 * (n + 2) + 16.125 * d * n for e in nnz)
    return fwd_layer, bwd_layer, L * (fwd_layer + bwd_layer)


_SAMPLER_SRC = r"""
import sys, time
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
print("max", mx, flush=True)
while True:
    print(time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), int(get(h)), flush=True)
    time.sleep(0.001)
"""


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: a helper process polls NVML
    every ~1 ms (a thread would starve behind the launch loop's GIL); samples are cut to the
    timed window by wall-clock time."""
    REASONS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.idx, self.proc, self.t0, self.t1 = gpu_index, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.5)          # let it import and start polling before the window opens
        except Exception:
            self.proc = None

    def open_window(self):
        self.t0 = time.time()

    def close_window(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.01)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=3)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.splitlines():
            f = line.split()
            try:
                if f[0] == "max":
                    mx = float(f[1])
                    continue
                t, clk, r = float(f[0]), float(f[1]), int(f[2])
            except Exception:
                continue
            if self.t0 is not None and self.t0 <= t <= (self.t1 or t):
                sm.append(clk)
                for n, bit in self.REASONS.items():
                    if r & bit:
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml, 1 ms polling, timed window only"}


def make_workload(args, rank):
    from sagnn_b200 import data_handler as dh
    shape = dict(dh.SHAPES[args.workload])
    L = args.layers or shape["L"]
    d = args.latdim or shape["d"]
    g = dh.make_named(args.workload, seed=args.seed + 1000 * rank, scale=args.scale)
    return g, L, d


def cpu_reference_run(g, L, d, steps, warmup, seed):
    """The TF1-graph restatement on the host cores; returns (seconds per step, threads)."""
    import torch
    from oracle import propagate_oracle as po, tf1_mirror
    from sagnn_b200 import data_handler as dh
    # all the host threads this process may use (torchrun presets OMP_NUM_THREADS=1 for its workers)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    T, U, I = g.graph_num, g.n_user, g.n_item
    adj = [torch.from_numpy(po.trans_to_lsts(m)[0].astype(np.int64)) for m in g.sub_mat]
    tp = [torch.from_numpy(po.trans_to_lsts(po.transpose(m))[0].astype(np.int64)) for m in g.sub_mat]
    uE = torch.from_numpy(dh.xavier_embeddings(T, U, d, seed))
    iE = torch.from_numpy(dh.xavier_embeddings(T, I, d, seed + 1))
    gen = torch.Generator().manual_seed(seed)
    gU = torch.randn((T, U, d), generator=gen)
    gI = torch.randn((T, I, d), generator=gen)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        tf1_mirror.propagate(adj, tp, uE, iE, gU, gI, L, 0.5)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return float(np.mean(times)), int(torch.get_num_threads())


def cpu_extra_baselines(g, L, d, seed):
    """SURVEY 8(d) b2 / b3 next to the stated baseline: the same TF1-graph restatement on ONE thread (one timed
    step) and a scipy CSR @ dense forward-only sanity floor (one thread, fp32, residual + layer sum, no backward)."""
    import torch
    from oracle import propagate_oracle as po, tf1_mirror
    from sagnn_b200 import data_handler as dh
    T, U, I = g.graph_num, g.n_user, g.n_item
    trav = 4 * L * sum(g.nnz)
    out = {}
    adj = [torch.from_numpy(po.trans_to_lsts(m)[0].astype(np.int64)) for m in g.sub_mat]
    tp = [torch.from_numpy(po.trans_to_lsts(po.transpose(m))[0].astype(np.int64)) for m in g.sub_mat]
    uE, iE = dh.xavier_embeddings(T, U, d, seed), dh.xavier_embeddings(T, I, d, seed + 1)
    gen = torch.Generator().manual_seed(seed)
    gU, gI = torch.randn((T, U, d), generator=gen), torch.randn((T, I, d), generator=gen)
    n_before = torch.get_num_threads()
    try:
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        tf1_mirror.propagate(adj, tp, torch.from_numpy(uE), torch.from_numpy(iE), gU, gI, L, 0.5)
        dt = time.perf_counter() - t0
        out["tf1_mirror_single_thread"] = {"value": trav / dt, "cores": 1, "ms_per_step": dt * 1e3,
                                           "sample": "1 timed fwd+bwd step, no warm-up"}
    finally:
        torch.set_num_threads(n_before)
    # forward only: binary structure @ dense table, LeakyReLU, residual, layer sum (model.py:118-127)
    import scipy.sparse as sp
    A = [sp.csr_matrix((np.ones(m.nnz, np.float32), m.indices, m.indptr), shape=m.shape) for m in g.sub_mat]
    At = [sp.csr_matrix(a.T) for a in A]
    t0 = time.perf_counter()
    for k in range(T):
        e0, e1 = uE[k], iE[k]
        s0, s1 = e0.copy(), e1.copy()
        for _ in range(L):
            z0, z1 = A[k] @ e1, At[k] @ e0
            e0, e1 = e0 + np.maximum(0.5 * z0, z0), e1 + np.maximum(0.5 * z1, z1)
            s0 += e0; s1 += e1
    dt = time.perf_counter() - t0
    out["scipy_csr_forward_only"] = {"value": 2 * L * sum(g.nnz) / dt, "cores": 1, "ms_per_forward": dt * 1e3,
                                     "note": "forward traversals only (2*L*E per pass), fp32, transposes prebuilt"}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        g, L, d = make_workload(args, 0)
        edges = sum(g.nnz)
        trav = 4 * L * edges
        sec, threads = cpu_reference_run(g, L, d, args.steps, max(1, args.warmup), args.seed)
        val = trav / sec
        line = {
            "impl": "reference", "metric": "fwd+bwd interval-graph SpMM edge traversals/s",
            "value": val, "unit": "edge_traversals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, g, L, d, world),
            "cpu_baseline": {"value": val, "unit": "edge_traversals/s", "cores": threads, "kind": "port",
                             "sample": "full %s workload, %d timed fwd+bwd steps of oracle/tf1_mirror.py "
                                       "(torch-CPU op-by-op restatement of the TF1 graph; TF 1.14 itself is "
                                       "not installable)" % (args.workload, args.steps)},
            "e2e": {"value": val, "unit": "edge_traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ own arm
    import torch
    import sagnn_b200 as sg
    from sagnn_b200 import data_handler as dh
    from sagnn_b200.step import PropagationStep

    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device; sagnn_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        # a rank that dies must not leave the others waiting in a collective for the default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    g, L, d = make_workload(args, rank)
    T, U, I = g.graph_num, g.n_user, g.n_item
    edges = sum(g.nnz)
    trav = 4 * L * edges

    t0 = time.perf_counter()
    plan = sg.build_plan(g.sub_mat, device=dev, latdim=d)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3

    if args.no_allgather:
        args.exchange = "none"
    auto_exchange = args.exchange == "auto"
    if auto_exchange:
        # N > 1: hand-off into symmetric-memory receive buffers -- written by the epilogue itself (peer stores) up to
        # 3 ranks, by the copy engines during the backward from 4 on (measured on 8 x B200: 0.550 vs 0.584 ms at N=8,
        # 0.563 vs 0.536 ms at N=2); falls back to the NCCL all-to-all if symmetric memory is unavailable
        args.exchange = "none" if world == 1 else ("fused" if world < 4 else "ce")
    if args.layout is None:
        args.layout = "rtd" if (world > 1 and args.exchange in ("alltoall", "fused", "ce")) else "trd"
    step = PropagationStep(plan, L, d, 0.5, layout=args.layout, row_multiple=world)
    step.u_embed.copy_(torch.from_numpy(dh.xavier_embeddings(T, U, d, args.seed)))
    step.i_embed.copy_(torch.from_numpy(dh.xavier_embeddings(T, I, d, args.seed + 1)))
    gen = torch.Generator(device=dev).manual_seed(args.seed + rank)
    step.g_user.normal_(generator=gen)
    step.g_item.normal_(generator=gen)

    do_gather = world > 1 and args.exchange != "none"
    zero_copy = do_gather and args.exchange == "alltoall" and args.layout == "rtd"
    if do_gather:
        if args.exchange == "allgather":
            gat_u = torch.empty((world,) + tuple(step.user_out.shape), dtype=torch.float32, device=dev)
            gat_i = torch.empty((world,) + tuple(step.item_out.shape), dtype=torch.float32, device=dev)
        elif zero_copy:       # my row block of every rank's intervals: [world, block, T, d]
            rcv_u = torch.empty((world, step.user_out_full.shape[0] // world, T, d), dtype=torch.float32, device=dev)
            rcv_i = torch.empty((world, step.item_out_full.shape[0] // world, T, d), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream()
    ce = do_gather and args.exchange == "ce"
    fused = do_gather and args.exchange in ("fused", "ce")      # both fill symmetric-memory receive buffers
    fused_ok = None
    if fused:
        # receive buffers in symmetric memory: every rank maps every peer's buffer, the forward's
        # epilogue stores each finished row straight into the consumer rank's buffer over NVLink
        from sagnn_b200.dist import exchange_rows_rtd
        bu, bi = step.user_out_full.shape[0] // world, step.item_out_full.shape[0] // world
        try:
            import torch.distributed._symmetric_memory as symm
            rcv_u = symm.empty((world, bu, T, d), dtype=torch.float32, device=dev)
            rcv_i = symm.empty((world, bi, T, d), dtype=torch.float32, device=dev)
            rcv_u.zero_(); rcv_i.zero_()
            hdl_u = symm.rendezvous(rcv_u, dist.group.WORLD)
            hdl_i = symm.rendezvous(rcv_i, dist.group.WORLD)
            have = 1
        except Exception as e:   # no symmetric memory on this box
            sys.stderr.write("[bench] rank %d: symmetric memory unavailable (%s)\n" % (rank, e))
            have = 0
        flag = torch.tensor([have], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not flag.item():
            if not auto_exchange:
                raise RuntimeError("--exchange fused needs torch symmetric memory on every rank")
            fused, zero_copy, args.exchange = False, True, "alltoall"
            rcv_u = torch.empty((world, bu, T, d), dtype=torch.float32, device=dev)
            rcv_i = torch.empty((world, bi, T, d), dtype=torch.float32, device=dev)
    if fused:
        # one-time check against the NCCL hand-off of the plain forward: same bits in every slab
        step.forward()
        ref_u, ref_i = exchange_rows_rtd(step.user_out_full), exchange_rows_rtd(step.item_out_full)
        step.set_scatter(world, rank, list(hdl_u.buffer_ptrs), list(hdl_i.buffer_ptrs))
        step.forward()
        hdl_u.barrier(channel=0)
        torch.cuda.synchronize()
        ok = torch.tensor([int(torch.equal(ref_u, rcv_u) and torch.equal(ref_i, rcv_i))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        fused_ok = bool(ok.item())
        if not fused_ok:
            raise RuntimeError("fused hand-off differs from the NCCL all-to-all of the same outputs")
        if ce:
            # copy-engine variant: the forward writes its [R,T,d] output locally, then one peer memcpy per row
            # block (DMA over NVLink, no SM) puts block j into rank j's receive buffer while the backward runs
            step.set_scatter(0, 0, None, None)
            peer_u = [hdl_u.get_buffer(r, (world, bu, T, d), torch.float32) for r in range(world)]
            peer_i = [hdl_i.get_buffer(r, (world, bi, T, d), torch.float32) for r in range(world)]
            rcv_u.zero_(); rcv_i.zero_()
            hdl_u.barrier(channel=0)

            def ce_step(with_barrier=False):
                step.forward()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for j in range(world):
                        r = (rank + j) % world               # start with my own block, then round the ring
                        peer_u[r][rank].copy_(step.user_out_full[r * bu:(r + 1) * bu], non_blocking=True)
                        peer_i[r][rank].copy_(step.item_out_full[r * bi:(r + 1) * bi], non_blocking=True)
                    if with_barrier:                         # arrival barrier behind the copies, under the backward
                        hdl_u.barrier(channel=0)
                step.backward()
                torch.cuda.current_stream().wait_stream(side)

            ce_step()
            hdl_u.barrier(channel=0)
            torch.cuda.synchronize()
            ok = torch.tensor([int(torch.equal(ref_u, rcv_u) and torch.equal(ref_i, rcv_i))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            fused_ok = bool(ok.item())
            if not fused_ok:
                raise RuntimeError("copy-engine hand-off differs from the NCCL all-to-all of the same outputs")

    split = None
    if not args.no_calibrate:
        split = step.calibrate(rounds=2)        # measured load balance (setup, like the plan build)
    # one CUDA graph per rank: the plain step, or forward-with-fused-hand-off + backward (peer pointers are
    # ordinary kernel arguments); the NCCL hand-offs keep direct launches
    use_graph = not args.no_graph and (not do_gather or fused)
    ce_graph = None
    barrier_in_graph = False

    def fused_step(with_barrier=False):     # forward with the epilogue's peer stores, arrival barrier under the backward
        step.forward()
        if with_barrier:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                hdl_u.barrier(channel=0)
        step.backward()
        if with_barrier:
            torch.cuda.current_stream().wait_stream(side)

    if use_graph and fused:
        # forward + hand-off + backward in ONE graph per rank; the symmetric-memory arrival barrier rides a side
        # stream inside it (behind the copies / the forward's peer stores, concurrent with the backward), so a step
        # no longer ends with a cross-rank rendezvous.  If the barrier cannot be captured it stays behind the graph.
        body = ce_step if ce else fused_step
        for in_graph in ((True, False) if os.environ.get("SAGNN_BARRIER_IN_GRAPH", "1") != "0" else (False,)):
            try:
                cap = torch.cuda.Stream()
                cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cap):
                    body(in_graph)
                    if not in_graph:
                        hdl_u.barrier(channel=0)
                torch.cuda.current_stream().wait_stream(cap)
                torch.cuda.synchronize()
                ce_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ce_graph):
                    body(in_graph)
                barrier_in_graph = in_graph
                break
            except Exception as e:
                if not in_graph:
                    raise
                sys.stderr.write("[bench] rank %d: barrier not capturable (%s), keeping it behind the graph\n" % (rank, e))
                ce_graph = None
                torch.cuda.synchronize()
        ok = torch.tensor([int(barrier_in_graph)], device=dev)      # every rank must run the same arrangement
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if barrier_in_graph and not bool(ok.item()):
            raise RuntimeError("ranks disagree on the barrier arrangement")
    elif use_graph:
        step.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    if args.no_flush:
        class _NoFlush:
            def zero_(self):
                pass
        flush = _NoFlush()

    def one_step():
        if fused and ce_graph is not None:
            ce_graph.replay()                                       # rows land in the peers' buffers as they finish / by DMA
            if not barrier_in_graph:
                hdl_u.barrier(channel=0)                            # every rank is through: my receive slabs are complete
        elif ce:
            ce_step()
            hdl_u.barrier(channel=0)
        elif fused:
            step.forward()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                hdl_u.barrier(channel=0)
            step.backward()
            torch.cuda.current_stream().wait_stream(side)
        elif do_gather:
            step.forward()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                           # overlaps the backward
                if args.exchange == "allgather":
                    dist.all_gather_into_tensor(gat_u, step.user_out)
                    dist.all_gather_into_tensor(gat_i, step.item_out)
                elif zero_copy:
                    from sagnn_b200.dist import exchange_rows_rtd
                    exchange_rows_rtd(step.user_out_full, rcv_u)
                    exchange_rows_rtd(step.item_out_full, rcv_i)
                else:
                    from sagnn_b200.dist import exchange_rows
                    exchange_rows(step.user_out)
                    exchange_rows(step.item_out)
            step.backward()
            torch.cuda.current_stream().wait_stream(side)
        else:
            step.replay()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.open_window()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record()
        one_step()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.close_window()
    clocks = sampler.stop() if rank == 0 else None
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(per_step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        e_all = torch.tensor([float(edges)], device=dev, dtype=torch.float64)
        dist.all_reduce(e_all, op=dist.ReduceOp.SUM)
        edges_all = float(e_all.item())
    else:
        edges_all = float(edges)
    ms_per_step = total_ms / args.steps
    value = 4 * L * edges_all / (ms_per_step * 1e-3)

    # ---- per-kernel durations for the roofline (direct launches, events on the launch stream)
    fwd_ms, bwd_ms = [], []
    for _ in range(max(5, min(args.steps, 20))):
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); step.forward(); b.record(); step.backward(); c.record()
        torch.cuda.synchronize()
        fwd_ms.append(a.elapsed_time(b)); bwd_ms.append(b.elapsed_time(c))
    fwd_ms, bwd_ms = float(np.mean(fwd_ms)), float(np.mean(bwd_ms))
    peak, peak_src = load_peaks()
    b_fwd_layer, b_bwd_layer, b_step = alg_bytes(g.nnz, U, I, d, L)
    dom_is_fwd = fwd_ms >= bwd_ms
    dom_ms = (fwd_ms if dom_is_fwd else bwd_ms) / L
    dom_bytes = b_fwd_layer if dom_is_fwd else b_bwd_layer
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")    # dram bytes per launch from an ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("fwd" if dom_is_fwd else "bwd")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of this command, not this run)"
        if traffic is not None else None, "peak_source": peak_src,
        "kernel": "%s<%s> (one GNN layer, all T intervals, both orientations)"
                  % ("spmm_rpw_kernel" if os.environ.get("SAGNN_KERNEL", "").lower().startswith("v8") else "spmm_pkt_kernel",
                     "FWD" if dom_is_fwd else "BWD"),
        "alg_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
        "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
        "step": {"alg_bytes": b_step, "achieved": b_step / (ms_per_step * 1e-3) / 1e9,
                 "frac": b_step / (ms_per_step * 1e-3) / 1e9 / peak,
                 "frac_of_nominal_8TBs": b_step / (ms_per_step * 1e-3) / 1e9 / 8000.0},
    }

    # ---- end to end through the C-ABI host entry point (pinned host buffers)
    e2e = None
    if not args.no_e2e:
        host = {}
        # the host entry point speaks the reference's [T,R,d] layout whatever layout the device step uses
        trd = lambda t: t if args.layout == "trd" else t.transpose(0, 1)
        # pinned pages on the NUMA node the GPU hangs off (first touch while bound to its cores; affinity restored after)
        numa = {}
        from sagnn_b200 import hostmem
        with (hostmem.near_gpu(dev.index, numa) if os.environ.get("SAGNN_NUMA_BIND", "1") != "0" else contextlib.nullcontext()):
            for name, src in (("uE", step.u_embed), ("iE", step.i_embed), ("gU", trd(step.g_user)), ("gI", trd(step.g_item))):
                host[name] = src.contiguous().cpu().pin_memory()
            for name, rows in (("uO", U), ("iO", I), ("dU", U), ("dI", I)):
                host[name] = torch.empty((T, rows, d), dtype=torch.float32).pin_memory()
                host[name].zero_()
        h2d = sum(host[k].numel() * 4 for k in ("uE", "iE", "gU", "gI"))
        d2h = sum(host[k].numel() * 4 for k in ("uO", "iO", "dU", "dI"))

        def host_step():
            sg.propagate_host(plan, host["uE"], host["iE"], host["gU"], host["gI"], host["uO"], host["iO"],
                              host["dU"], host["dI"], L, 0.5)
        for _ in range(3):
            host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_step()                    # synchronous: returns when the results are in host memory
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        chk = float(host["uO"][0, 0, 0])   # the device->host result is really there
        e2e = {"value": 4 * L * edges_all / e2e_s, "unit": "edge_traversals/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "probe": chk,
               "api": "sagnn_propagate_host (C ABI, pinned host buffers)",
               "pinned_pages_numa_local": bool(numa.get("bound", False))}

    # ---- CPU baseline next to it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, threads = cpu_reference_run(g, L, d, args.cpu_steps, 1, args.seed)
        cpu = {"value": trav / sec, "unit": "edge_traversals/s", "cores": threads, "kind": "port",
               "ms_per_step": sec * 1e3,
               "sample": "full %s workload, %d timed fwd+bwd steps of oracle/tf1_mirror.py (torch-CPU "
                         "op-by-op restatement of the TF1 graph)" % (args.workload, args.cpu_steps)}
        try:
            from oracle import c_oracle, propagate_oracle as po
            adj = [po.trans_to_lsts(m)[0] for m in g.sub_mat]
            tpl = [po.trans_to_lsts(po.transpose(m))[0] for m in g.sub_mat]
            arrs = [step.u_embed.cpu().numpy(), step.i_embed.cpu().numpy(), step.g_user.cpu().numpy(),
                    step.g_item.cpu().numpy()]
            c_oracle.propagate(adj, tpl, *arrs, L, 0.5, np.float32)
            t0 = time.perf_counter()
            c_oracle.propagate(adj, tpl, *arrs, L, 0.5, np.float32)
            cs = time.perf_counter() - t0
            cpu["fused_c_port"] = {"value": trav / cs, "cores": c_oracle.num_threads(),
                                   "note": "oracle/csrc/sagnn_oracle.c (fused OpenMP fp32, incl. CSR build)"}
        except Exception as e:   # the C port is optional colour, never fatal
            cpu["fused_c_port"] = {"error": str(e)}
        try:                     # single-thread restatement, scipy forward-only floor
            cpu["others"] = cpu_extra_baselines(g, L, d, args.seed)
        except Exception as e:   # optional colour as well
            cpu["others"] = {"error": str(e)}

    # ---- the configurations north_star names for N GPUs (results only; the headline above is untouched)
    extras = None
    launches = step.kernel_launches_per_step * args.steps
    stats = plan.stats()
    if args.extras == "on" or (args.extras == "auto" and world > 1):
        import bench_multi
        del flush
        step = None
        torch.cuda.empty_cache()
        extras = bench_multi.run_extras(args, rank, world, dev, dist, peak)

    if rank == 0:
        line = {
            "metric": "fwd+bwd interval-graph SpMM edge traversals/s", "value": value,
            "unit": "edge_traversals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, g, L, d, world),
            "run": dict(run_config(args, world, stats, use_graph, do_gather), cta_split=split,
                        **({"arrival_barrier": "inside the graph, side stream, concurrent with the backward" if barrier_in_graph
                            else "one launch behind the graph"} if fused else {})),
            "graph_edges_per_s": edges_all / (ms_per_step * 1e-3),
            "plan_build_ms": plan_ms, "wall_s_timed_region": wall,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if fused_ok is not None:
            line["run"]["fused_handoff_bitwise_equal_to_nccl_alltoall"] = fused_ok
        if cpu is not None and "fused_c_port" in cpu and "value" in cpu["fused_c_port"]:
            # both CPU arms next to each other: the stated baseline (TF1-graph restatement) and the fused C port
            cpu["ratios"] = {"device_resident_vs_tf1_mirror": value / cpu["value"],
                             "device_resident_vs_fused_c_port": value / cpu["fused_c_port"]["value"],
                             "e2e_vs_tf1_mirror": (e2e["value"] / cpu["value"]) if e2e else None,
                             "e2e_vs_fused_c_port": (e2e["value"] / cpu["fused_c_port"]["value"]) if e2e else None}
        line["extras"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def workload_config(args, g, L, d, world):
    """What is measured (identical for both arms); how it was run goes into ``run``."""
    return {
        "workload": "%s-shaped synthetic power-law interval graphs (SURVEY app. D)" % args.workload,
        "T": g.graph_num, "U": g.n_user, "I": g.n_item, "interval_edges": g.nnz, "edges": sum(g.nnz),
        "layers": L, "latdim": d, "leaky": 0.5, "scale": args.scale, "seed": args.seed,
        "edge_traversals_per_step": 4 * L * sum(g.nnz),
        "per_gpu": "every rank owns one full set of T interval graphs (interval sharding, weak scaling)"
                   if world > 1 else "single GPU",
    }


def run_config(args, world, stats, use_graph, gather):
    cfg = {"l2": "256 MB buffer written between timed steps (L2 flush), outside the event pairs",
           "schedule": stats, "cuda_graph": bool(use_graph),
           "kernel": os.environ.get("SAGNN_KERNEL", "v10 packet stream (default)"),
           "layout": "[T,R,d] outputs (model.py:131-132)" if args.layout == "trd"
                     else "[R,T,d] outputs / upstream (model.py:133-134, fused transpose)",
           "allgather_outputs": bool(gather) and args.exchange == "allgather",
           "exchange": args.exchange if world > 1 else "none"}
    if world > 1:
        cfg["exchange_detail"] = {
            "alltoall": "one NCCL all-to-all per side: every rank receives its row block of all ranks' intervals "
                        "(row-sharded consumer), sent straight from the [R,T,d] epilogue output, on a side "
                        "stream overlapping the backward",
            "allgather": "NCCL all-gather of the [T,R,d] outputs to every rank (replicated consumer), on a side "
                         "stream overlapping the backward",
            "fused": "no collective kernel: the last forward layer's epilogue stores every finished row into the "
                     "symmetric-memory receive buffer of the rank that owns its row block (peer stores over "
                     "NVLink, sagnn_propagate_fwd_scatter); forward + arrival barrier + backward replay from one CUDA graph "
                     "(the symmetric-memory barrier on a side stream behind the forward, see arrival_barrier); verified "
                     "bitwise against the NCCL all-to-all before timing",
            "ce": "no collective kernel and no SM: after the forward, one peer memcpy per row block (copy engines over "
                  "NVLink) fills the consumers' symmetric-memory receive buffers on a side stream while the backward runs; "
                  "forward + copies + arrival barrier + backward replay from one CUDA graph (the symmetric-memory barrier "
                  "on the side stream behind the copies, see arrival_barrier); verified bitwise against the NCCL "
                  "all-to-all before timing",
            "none": "no hand-off collective (compute only)"}[args.exchange]
    return cfg


if __name__ == "__main__":
    sys.exit(main())
