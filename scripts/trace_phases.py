"""Per-phase cycle totals of the layer kernel (debug build with -DSAGNN_PHASES): scripts/trace_phases.py"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sagnn_b200 as sg
from sagnn_b200 import data_handler as dh, _lib
from sagnn_b200.step import PropagationStep
g = dh.make_named("gowalla", seed=100)
L, d = 2, 64
plan = sg.build_plan(g.sub_mat)
step = PropagationStep(plan, L, d)
step.u_embed.copy_(torch.from_numpy(dh.xavier_embeddings(3, g.n_user, d, 100)))
step.i_embed.copy_(torch.from_numpy(dh.xavier_embeddings(3, g.n_item, d, 101)))
step.g_user.normal_(); step.g_item.normal_()
for _ in range(3): step.run()
torch.cuda.synchronize()
sms = plan.stats()["sms"]; W = 32
per = sms * 4 + sms * W * 4
buf = torch.zeros(4 * per, dtype=torch.int64, device="cuda")
lib = _lib.load_library()
# capacity is counted in units of sms*4 words by the library: give it launches*(1+W)
_lib.check(lib.sagnn_debug_trace(plan.handle, ctypes.c_void_p(buf.data_ptr()), 1))
step.forward(); torch.cuda.synchronize()
t = buf.cpu().numpy()
names = ["top (record, queue, own-row issue)", "gather", "long-row publish/reduce", "epilogue"]
for l in range(1):
    ph = t[sms * 4: sms * 4 + sms * W * 4].reshape(sms * W, 4).astype(np.float64)
    tot = ph.sum(1)
    print("launch 0: mean cycles per warp %.0f (min %.0f max %.0f)" % (tot.mean(), tot.min(), tot.max()))
    for i, n in enumerate(names):
        print("   %-40s %5.1f %%   mean %.0f cycles/warp" % (n, 100 * ph[:, i].sum() / ph.sum(), ph[:, i].mean()))
