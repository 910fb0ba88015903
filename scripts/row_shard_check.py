"""torchrun check + timing of row sharding over NCCL (second way).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/row_shard_check.py [--workload gowalla --scale 1.0 --steps 10]

Every rank builds a row-block plan of the SAME T graphs, runs RowShardedPropagation (per-layer table
all-gathers over NVLink) and compares outputs and gradients bitwise with the single-GPU
``propagate`` of the whole graph (rank 0 prints ROW_SHARD_OK).  With --steps it also times the
fwd+bwd step on the device (CUDA events, max over ranks) next to the single-GPU step.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sagnn_b200 as sg                      # noqa: E402
from sagnn_b200 import data_handler as dh    # noqa: E402
from sagnn_b200 import dist as sd            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="small")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--latdim", type=int, default=64)
    ap.add_argument("--steps", type=int, default=0)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    g = dh.make_named(args.workload, seed=100, scale=args.scale)
    T, U, I, L, d = g.graph_num, g.n_user, g.n_item, args.layers, args.latdim
    gen = torch.Generator(device=dev).manual_seed(7)            # same tables on every rank
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
    ref = [uv1.detach(), iv1.detach(), uE.grad, iE.grad]
    ok = all(torch.equal(a, b) for a, b in zip(got, ref))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res = {"world": world, "workload": args.workload, "T": T, "U": U, "I": I, "edges": int(sum(g.nnz)),
           "layers": L, "latdim": d, "bitwise_equal_to_single_gpu": bool(flag.item())}

    if args.steps:
        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def step_sharded():
            uE.grad = iE.grad = None
            o = rs(uE, iE)
            torch.autograd.backward(list(o), [gU, gI])

        def step_single():
            uE.grad = iE.grad = None
            o = sg.propagate(plan, uE, iE, L, 0.5)
            torch.autograd.backward(list(o), [gU, gI])

        res["ms_row_sharded"] = timed(step_sharded)
        res["ms_single_gpu"] = timed(step_single)
        res["edge_traversals_per_s_row_sharded"] = 4 * L * res["edges"] / (res["ms_row_sharded"] * 1e-3)
    if rank == 0:
        print(json.dumps(res))
        print("ROW_SHARD_OK" if res["bitwise_equal_to_single_gpu"] else "ROW_SHARD_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if res["bitwise_equal_to_single_gpu"] else 1)


if __name__ == "__main__":
    main()
