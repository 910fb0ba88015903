"""Pre-allocated fwd+bwd step over a plan (what a training loop calls once per batch; the
reference re-runs the full-graph propagation on every ``sess.run``, model.py:373,459).

All buffers are allocated once; ``run()`` only enqueues the C-ABI calls on the current stream,
so the whole step can be captured into a CUDA graph (``capture()``) and replayed with no
per-layer launch latency.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from .propagate import _ptr, _stream_ptr


class PropagationStep:
    def __init__(self, plan, n_layers, d, leaky=0.5, layout="trd", row_multiple=1):
        """layout="rtd": outputs and upstream gradients are [R,T,d] (model.py:133-134, the LSTM's
        input layout; SAGNN_LAYOUT_RTD), embeddings and their gradients stay [T,R,d].
        row_multiple (rtd only): the output buffers ``user_out_full`` / ``item_out_full`` get zero rows
        up to a multiple of it, so that equal row blocks can be sent to a row-sharded consumer straight
        from the epilogue's output (``dist.exchange_rows_rtd``); ``user_out`` / ``item_out`` are the
        [:R] prefixes the kernel writes."""
        self.plan, self.L, self.d, self.leaky = plan, int(n_layers), int(d), float(leaky)
        self.flags = {"trd": 0, "rtd": 1}[layout]
        dev = plan.device
        self.lib = _lib.load_library()
        with torch.cuda.device(dev):
            f, m, b = plan.workspace_bytes(self.L, self.d)
            self.ws = torch.empty(max(f, b, 1), dtype=torch.uint8, device=dev)
            self.masks = torch.empty(max(m, 1), dtype=torch.uint8, device=dev)
            mk = lambda rows: torch.empty((plan.T, rows, self.d), dtype=torch.float32, device=dev)
            self.u_embed, self.i_embed = mk(plan.U), mk(plan.I)
            mo = mk if not self.flags else \
                (lambda rows: torch.empty((rows, plan.T, self.d), dtype=torch.float32, device=dev))
            self.g_user, self.g_item = mo(plan.U), mo(plan.I)
            if self.flags and row_multiple > 1:
                pad = lambda rows: -(-rows // row_multiple) * row_multiple
                self.user_out_full = torch.zeros((pad(plan.U), plan.T, self.d), dtype=torch.float32, device=dev)
                self.item_out_full = torch.zeros((pad(plan.I), plan.T, self.d), dtype=torch.float32, device=dev)
                self.user_out, self.item_out = self.user_out_full[:plan.U], self.item_out_full[:plan.I]
            else:
                self.user_out, self.item_out = mo(plan.U), mo(plan.I)
                self.user_out_full, self.item_out_full = self.user_out, self.item_out
            self.d_u, self.d_i = mk(plan.U), mk(plan.I)
        self.graph = None
        # L forward + L backward layer kernels + the streaming pre-mask of the upstream
        self.kernel_launches_per_step = 2 * self.L + 1

    def set_scatter(self, world, rank, user_ptrs, item_ptrs):
        """Fused hand-off (``sagnn_propagate_fwd_scatter``, layout "rtd" only): from now on ``forward()``
        writes the finished layer sums straight into the ranks' receive buffers ``[world, blk, T, d]``
        (``user_ptrs`` / ``item_ptrs``: their device pointers as mapped on THIS device, own buffer
        included) instead of ``user_out`` / ``item_out``.  ``set_scatter(0, 0, None, None)`` switches back."""
        import ctypes
        if not world:
            self.scatter = None
            return
        if not self.flags:
            raise ValueError("the fused hand-off needs layout='rtd'")
        arr = lambda ptrs: (ctypes.c_void_p * world)(*[int(x) for x in ptrs])
        self.scatter = (int(world), int(rank), arr(user_ptrs), arr(item_ptrs))

    def forward(self):
        p = self.plan
        if getattr(self, "scatter", None):
            w, r, up, ip = self.scatter
            _lib.check(self.lib.sagnn_propagate_fwd_scatter(
                p.handle, _ptr(self.u_embed), _ptr(self.i_embed), _ptr(self.user_out), _ptr(self.item_out), self.L,
                self.d, self.leaky, _ptr(self.masks), _ptr(self.ws), self.ws.numel(), w, r, up, ip,
                _stream_ptr(p.device)))
            return
        _lib.check(self.lib.sagnn_propagate_fwd_ex(p.handle, _ptr(self.u_embed), _ptr(self.i_embed),
                                                   _ptr(self.user_out), _ptr(self.item_out), self.L, self.d,
                                                   self.leaky, _ptr(self.masks), _ptr(self.ws), self.ws.numel(),
                                                   self.flags, _stream_ptr(p.device)))

    def backward(self):
        p = self.plan
        _lib.check(self.lib.sagnn_propagate_bwd_ex(p.handle, _ptr(self.g_user), _ptr(self.g_item), _ptr(self.d_u),
                                                   _ptr(self.d_i), self.L, self.d, self.leaky, _ptr(self.masks),
                                                   _ptr(self.ws), self.ws.numel(), self.flags,
                                                   _stream_ptr(p.device)))

    def forward_interval(self, k):
        """Forward restricted to interval k (all SMs on its two CSRs); see sagnn_propagate_fwd_interval."""
        p = self.plan
        _lib.check(self.lib.sagnn_propagate_fwd_interval(p.handle, int(k), _ptr(self.u_embed), _ptr(self.i_embed),
                                                         _ptr(self.user_out), _ptr(self.item_out), self.L, self.d,
                                                         self.leaky, _ptr(self.masks), _ptr(self.ws),
                                                         self.ws.numel(), _stream_ptr(p.device)))

    def backward_interval(self, k):
        p = self.plan
        _lib.check(self.lib.sagnn_propagate_bwd_interval(p.handle, int(k), _ptr(self.g_user), _ptr(self.g_item),
                                                         _ptr(self.d_u), _ptr(self.d_i), self.L, self.d, self.leaky,
                                                         _ptr(self.masks), _ptr(self.ws), self.ws.numel(),
                                                         _stream_ptr(p.device)))

    def run(self):
        self.forward()
        self.backward()

    def calibrate(self, rounds=2):
        """Measured load balancing: trace one step per round (per-CTA timestamps from the kernels),
        take work(segment) = sum over the step's launches of mean CTA busy time x CTAs, and re-deal
        the persistent CTAs accordingly (``sagnn_plan_rebalance``).  Uses the step's current buffers;
        results are unaffected (only the CTA -> segment table changes).  Returns the final split."""
        import ctypes
        import numpy as np
        p = self.plan
        sms = p.stats()["sms"]
        S = 2 * p.T
        n_launch = 2 * self.L
        with torch.cuda.device(p.device):
            for _ in range(rounds):
                buf = torch.zeros(n_launch * sms * 4, dtype=torch.int64, device=p.device)
                _lib.check(self.lib.sagnn_debug_trace(p.handle, ctypes.c_void_p(buf.data_ptr()), n_launch))
                self.run()
                torch.cuda.synchronize(p.device)
                _lib.check(self.lib.sagnn_debug_trace(p.handle, None, 0))
                t = buf.cpu().numpy().reshape(n_launch, sms, 4)
                work = np.zeros(S)
                for l in range(n_launch):
                    seg, busy = t[l, :, 0], (t[l, :, 3] - t[l, :, 1]).astype(np.float64)
                    for s in range(S):
                        m = seg == s
                        if m.any():
                            work[s] += busy[m].mean() * m.sum()
                work = np.maximum(work, 1.0)
                arr = (ctypes.c_double * S)(*work.tolist())
                _lib.check(self.lib.sagnn_plan_rebalance(p.handle, arr, _stream_ptr(p.device)))
            out = (ctypes.c_int * S)()
            _lib.check(self.lib.sagnn_plan_get_split(p.handle, out))
        return list(out)

    def capture(self):
        """Captures fwd+bwd into a CUDA graph; afterwards ``replay()`` launches the whole step."""
        with torch.cuda.device(self.plan.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.run()                                   # warm-up outside capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run()
            self.graph = g
        return self

    def replay(self):
        if self.graph is None:
            self.run()
        else:
            self.graph.replay()
