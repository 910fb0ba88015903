"""A/B of the schedule order (SAGNN_SORT_WINDOW, csrc/plan.cu sched_key_kernel): one process, one plan per setting
(the variable is read at plan finalize), fwd+bwd step replayed from a CUDA graph, L2 flushed between steps, CUDA events.

    python scripts/sort_window_ab.py [--workloads gowalla,amazon-book] [--windows 0,1024,4096,16384] [--steps 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sagnn_b200 as sg                                      # noqa: E402
from sagnn_b200 import data_handler as dh                    # noqa: E402
from sagnn_b200.step import PropagationStep                  # noqa: E402


def time_step(step, steps, warmup, flush):
    ms = []
    for it in range(warmup + steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step.replay(); b.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ms.append(a.elapsed_time(b))
    return float(np.mean(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="gowalla")
    ap.add_argument("--windows", default="0,1024,4096,16384")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/sort_window_ab.json")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = []
    for wl in args.workloads.split(","):
        g = dh.make_named(wl, seed=100)
        shape = dh.SHAPES[wl]
        L, d = int(shape.get("L", shape.get("gnn_layer", 2))), int(shape.get("d", shape.get("latdim", 64)))
        ref = None
        for w in [int(x) for x in args.windows.split(",")]:
            os.environ["SAGNN_SORT_WINDOW"] = str(w)
            plan = sg.build_plan(g.sub_mat, latdim=d)
            step = PropagationStep(plan, L, d, 0.5)
            gen = torch.Generator(device="cuda").manual_seed(1)
            for t in (step.u_embed, step.i_embed, step.g_user, step.g_item):
                t.normal_(generator=gen)
            step.calibrate(rounds=2)
            step.capture()
            mean, best = time_step(step, args.steps, args.warmup, flush)
            outs = [x.clone() for x in (step.user_out, step.item_out, step.d_u, step.d_i)]
            same = True if ref is None else all(torch.equal(a, b) for a, b in zip(outs, ref))
            if ref is None:
                ref = outs
            r = dict(workload=wl, L=L, d=d, window=w, ms_mean=round(mean, 4), ms_min=round(best, 4), bitwise_equal_to_window0=same)
            print(json.dumps(r), flush=True)
            res.append(r)
            del step, plan
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
