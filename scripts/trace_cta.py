"""Per-CTA / per-segment timeline of the layer launches (diagnostics): python scripts/trace_cta.py [workload]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sagnn_b200 as sg
from sagnn_b200 import data_handler as dh, _lib
from sagnn_b200.step import PropagationStep

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "gowalla"
g = dh.make_named(name, seed=100)
L, d = g.meta["L"], g.meta["d"]
plan = sg.build_plan(g.sub_mat)
step = PropagationStep(plan, L, d)
step.u_embed.copy_(torch.from_numpy(dh.xavier_embeddings(g.graph_num, g.n_user, d, 100)))
step.i_embed.copy_(torch.from_numpy(dh.xavier_embeddings(g.graph_num, g.n_item, d, 101)))
step.g_user.normal_(); step.g_item.normal_()
if '--calibrate' in sys.argv:
    print('calibrated split', step.calibrate(rounds=2))
for _ in range(3):
    step.run()
torch.cuda.synchronize()
sms = plan.stats()["sms"]
n_launch = 2 * L
buf = torch.zeros(n_launch * sms * 4, dtype=torch.int64, device="cuda")
lib = _lib.load_library()
_lib.check(lib.sagnn_debug_trace(plan.handle, ctypes.c_void_p(buf.data_ptr()), n_launch))
step.run()
torch.cuda.synchronize()
_lib.check(lib.sagnn_debug_trace(plan.handle, None, 0))
t = buf.cpu().numpy().reshape(n_launch, sms, 4)
for l in range(n_launch):
    seg, t0, t1, t2 = t[l, :, 0], t[l, :, 1], t[l, :, 2], t[l, :, 3]
    base = t0.min()
    print("launch %d (%s): kernel span %.1f us, staging %.1f us (mean)" % (l, "fwd" if l < L else "bwd", (t2.max() - base) / 1e3, (t1 - t0).mean() / 1e3))
    for s in np.unique(seg):
        m = seg == s
        print("   seg %d (k=%d %s): %3d CTAs  end min/mean/max = %.1f / %.1f / %.1f us" % (
            s, s // 2, "item" if s % 2 else "user", m.sum(), (t2[m].min() - base) / 1e3, (t2[m].mean() - base) / 1e3, (t2[m].max() - base) / 1e3))
