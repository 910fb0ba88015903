"""Multi-GPU plumbing for the propagation path: interval sharding + output all-gather, and row
sharding of the graphs (one table all-gather per layer) for when there are fewer intervals than GPUs.

The T interval graphs share nothing -- separate adjacency, separate embedding slices
``uEmbed[k]`` / ``iEmbed[k]`` (model.py:108-109,119-120), separate outputs (model.py:128-129) --
so rank r owns a subset of the intervals, runs forward and backward with NO data-path
collective, and a single all-gather of the per-interval layer sums hands ``[T,U,d]`` /
``[T,I,d]`` to the (replicated) interval-fusion stage (model.py:131-155).  One process per GPU,
``torch.distributed`` (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def assign_intervals(nnz, world):
    """Longest-processing-time assignment of intervals to ranks by edge count.
    Returns ``owners`` (list of rank per interval).  Deterministic; with T <= world every
    interval gets its own rank."""
    order = sorted(range(len(nnz)), key=lambda k: (-int(nnz[k]), k))
    load = [0] * world
    owners = [0] * len(nnz)
    for k in order:
        r = min(range(world), key=lambda i: (load[i], i))
        owners[k] = r
        load[r] += int(nnz[k])
    return owners


def local_intervals(owners, rank):
    return [k for k, r in enumerate(owners) if r == rank]


class _GatherIntervals(torch.autograd.Function):
    """all-gather of per-interval tensors [T_local, R, d] -> [T, R, d] in interval order.
    Backward: the consumer is replicated (every rank holds the full upstream), so the
    gradient of the local shard is the local slice -- no collective in the backward."""

    @staticmethod
    def forward(ctx, local, owners, rank, group):
        world = dist.get_world_size(group)
        T = len(owners)
        counts = [sum(1 for o in owners if o == r) for r in range(world)]
        cap = max(counts)
        R, d = local.shape[1], local.shape[2]
        if cap == 0:
            raise ValueError("no intervals")
        padded = local
        if local.shape[0] < cap:
            padded = torch.cat([local, local.new_zeros((cap - local.shape[0], R, d))], dim=0)
        gathered = local.new_empty((world, cap, R, d))
        dist.all_gather_into_tensor(gathered.view(world * cap, R, d), padded.contiguous(), group=group)
        out = local.new_empty((T, R, d))
        slot = [0] * world
        for k, o in enumerate(owners):
            out[k] = gathered[o, slot[o]]
            slot[o] += 1
        ctx.mine = [k for k, o in enumerate(owners) if o == rank]
        return out

    @staticmethod
    def backward(ctx, g):
        idx = torch.tensor(ctx.mine, dtype=torch.long, device=g.device)
        return g.index_select(0, idx), None, None, None


def gather_intervals(local, owners, rank=None, group=None):
    rank = dist.get_rank(group) if rank is None else rank
    return _GatherIntervals.apply(local, owners, rank, group)


def row_blocks(n_rows, world):
    """Equal row blocks (the last one may be short): block b = rows [b*size, min((b+1)*size, n_rows))."""
    size = (n_rows + world - 1) // world
    return size, [(min(b * size, n_rows), min((b + 1) * size, n_rows)) for b in range(world)]


def exchange_rows(local, group=None):
    """Hand-off to a ROW-sharded consumer: instead of gathering every interval
    everywhere, rank j receives only its block of rows of every rank's intervals.

    local [T_local, R, d] (the same T_local on every rank) -> [world*T_local, block, d]: the
    caller's row block of all intervals, in rank order.  One all-to-all of 1/world the all-gather
    volume.  Rows are padded to a multiple of world (pad rows are zero)."""
    world = dist.get_world_size(group)
    T, R, d = local.shape
    size, _ = row_blocks(R, world)
    if size * world != R:
        local = torch.cat([local, local.new_zeros((T, size * world - R, d))], dim=1)
    send = local.view(T, world, size, d).permute(1, 0, 2, 3).contiguous()      # [dst, T, block, d]
    recv = torch.empty_like(send)                                               # [src, T, block, d]
    dist.all_to_all_single(recv.view(world, -1), send.view(world, -1), group=group)
    return recv.view(world * T, size, d)


def exchange_rows_rtd(out_full, recv=None, group=None):
    """The same hand-off with NO pack / pad copies, for outputs the epilogue already wrote in the
    ``[R_pad, T_local, d]`` layout (``SAGNN_LAYOUT_RTD``; ``R_pad`` a multiple of the world size, pad
    rows zero -- ``PropagationStep(layout="rtd", row_multiple=world).user_out_full``): the row block
    of rank j is the contiguous slab ``out_full[j*b:(j+1)*b]``, so the output buffer IS the send
    buffer of one all-to-all.  Returns ``[world, b, T_local, d]``: my row block of every rank's
    intervals (``recv``: optional pre-allocated result)."""
    world = dist.get_world_size(group)
    Rp, T, d = out_full.shape
    if Rp % world or not out_full.is_contiguous():
        raise ValueError("need a contiguous [R_pad, T, d] tensor with R_pad (%d) a multiple of the world size" % Rp)
    b = Rp // world
    if recv is None:
        recv = torch.empty((world, b, T, d), dtype=out_full.dtype, device=out_full.device)
    dist.all_to_all_single(recv.view(world, -1), out_full.view(world, -1), group=group)
    return recv


class ShardedPropagation:
    """Interval-sharded drop-in for ``propagate``: every rank passes the FULL parameter tables and gets
    the full ``[T,U,d]`` / ``[T,I,d]`` back.

    Gradient contract: the backward returns parameter gradients that are non-zero only in the intervals
    THIS rank owns (the consumer is replicated, so every rank already holds the full upstream and no
    collective runs in the backward).  With replicated ``uEmbed`` / ``iEmbed`` the caller must
    SUM-all-reduce those gradients across the group before the optimizer step (intervals are disjoint, so
    the sum just assembles them; DDP's mean would scale them by 1/world) -- or keep the parameters sharded
    by interval like the plan.  ``tests/test_dist_cpu.py`` pins this contract."""

    def __init__(self, sub_mats, U=None, I=None, n_layers=2, leaky=0.5, group=None, device=None,
                 edge_weight=None):
        from .propagate import build_plan
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        nnz = [int(m.nnz) if hasattr(m, "nnz") else int(len(m[0]) if isinstance(m, (tuple, list)) else len(m))
               for m in sub_mats]
        self.owners = assign_intervals(nnz, self.world)
        self.mine = local_intervals(self.owners, self.rank)
        self.n_layers, self.leaky = n_layers, leaky
        self.plan = None
        if self.mine:
            ew = [edge_weight[k] for k in self.mine] if isinstance(edge_weight, (list, tuple)) else edge_weight
            self.plan = build_plan([sub_mats[k] for k in self.mine], U, I, device=device, edge_weight=ew)

    def __call__(self, u_embed, i_embed):
        from .propagate import propagate
        idx = torch.tensor(self.mine, dtype=torch.long, device=u_embed.device)
        if self.plan is not None:
            uv, iv = propagate(self.plan, u_embed.index_select(0, idx), i_embed.index_select(0, idx),
                               self.n_layers, self.leaky)
        else:   # more ranks than intervals: this rank only takes part in the gather
            uv = u_embed.new_zeros((0,) + tuple(u_embed.shape[1:]))
            iv = i_embed.new_zeros((0,) + tuple(i_embed.shape[1:]))
        return (gather_intervals(uv, self.owners, self.rank, self.group),
                gather_intervals(iv, self.owners, self.rank, self.group))


# ---------------------------------------------------------------------------------------------
# Row sharding (second way): every rank holds all T graphs but computes only one block
# of user rows of each A_k and one block of item rows of each A_k^T.  A row gathers from every
# row of the other side, so each layer's freshly written table is all-gathered before the next
# layer reads it: L-1 table all-gathers in the forward, L in the backward (the pre-masked
# gradients), plus one for the outputs / parameter gradients when the consumer is replicated.
# Both orientations are stored, so there is never a reduce-scatter or a float atomic.
# ---------------------------------------------------------------------------------------------
def padded_rows(n_rows, world):
    """Rows rounded up so that every rank owns an equal block (pad rows are empty graph rows)."""
    size, _ = row_blocks(n_rows, world)
    return size * world


def forward_stages(n_layers):
    """[(layer, table index to exchange afterwards or None)] -- ``sagnn_propagate_fwd_layers``."""
    return [(l, l if l < n_layers - 1 else None) for l in range(n_layers)]


def backward_stages(n_layers):
    """[(phase, table index to exchange afterwards or None)] -- ``sagnn_propagate_bwd_levels``:
    phase 0 pre-masks the owned rows of the upstream, phase j >= 1 is level kernel j-1."""
    return [(ph, ph if ph < n_layers else None) for ph in range(n_layers + 1)]


def allgather_row_blocks(tables, rank=None, group=None):
    """In-place all-gather of equal row blocks of one table or a list of tables: each
    ``table [T, R, d]`` (R divisible by the world size) holds this rank's rows
    ``[rank*b, (rank+1)*b)`` of every interval on entry and all rows on exit.  One all-gather per
    interval and table, sent straight from / received straight into the table; on NCCL all of them
    are coalesced into ONE grouped launch (other backends: a loop with a staged send block)."""
    single = isinstance(tables, torch.Tensor)
    tabs = [tables] if single else list(tables)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group) if rank is None else rank
    for t in tabs:
        if t.shape[1] % world:
            raise ValueError("rows (%d) must be a multiple of the world size (%d)" % (t.shape[1], world))
    if world > 1:
        pairs = []
        for t in tabs:
            b = t.shape[1] // world
            pairs += [(t[k], t[k, rank * b:(rank + 1) * b]) for k in range(t.shape[0])]
        if tabs[0].is_cuda:
            # one NCCL launch for all (interval, side) pairs when torch offers its (private) coalescing manager;
            # otherwise the same in-place all-gathers one by one
            cm = getattr(dist, "_coalescing_manager", None)
            if cm is not None:
                with cm(group=group):
                    for out, mine in pairs:
                        dist.all_gather_into_tensor(out, mine, group=group)
            else:
                for out, mine in pairs:
                    dist.all_gather_into_tensor(out, mine, group=group)
        else:
            for out, mine in pairs:
                dist.all_gather_into_tensor(out, mine.clone(), group=group)
    return tables


class _RowShardedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_embed, i_embed, rs):
        be = rs._backend(u_embed.shape[2])
        u, i = rs._pad(u_embed, rs.U, rs.U_pad), rs._pad(i_embed, rs.I, rs.I_pad)
        need_bwd = u_embed.requires_grad or i_embed.requires_grad
        u_out, i_out = torch.empty_like(u), torch.empty_like(i)
        be.begin_forward(u, i, u_out, i_out, need_bwd)
        for l, ex in forward_stages(rs.n_layers):
            be.fwd_layers(l, l + 1)
            if ex is not None:
                allgather_row_blocks(be.table(0, ex), rs.rank, rs.group)
        allgather_row_blocks([u_out, i_out], rs.rank, rs.group)  # replicated consumer (model.py:131-155)
        ctx.rs, ctx.be, ctx.masks = rs, be, be.masks
        return u_out[:, :rs.U], i_out[:, :rs.I]

    @staticmethod
    def backward(ctx, g_user, g_item):
        rs, be = ctx.rs, ctx.be
        gu = rs._pad(g_user.contiguous(), rs.U, rs.U_pad)
        gi = rs._pad(g_item.contiguous(), rs.I, rs.I_pad)
        d_u, d_i = torch.empty_like(gu), torch.empty_like(gi)
        be.masks = ctx.masks                                    # the masks of THIS forward
        be.begin_backward(gu, gi, d_u, d_i)
        for ph, ex in backward_stages(rs.n_layers):
            be.bwd_levels(ph, ph + 1)
            if ex is not None:
                allgather_row_blocks(be.table(1, ex), rs.rank, rs.group)
        allgather_row_blocks([d_u, d_i], rs.rank, rs.group)      # replicated parameters
        return d_u[:, :rs.U], d_i[:, :rs.I], None


class _CudaRowBackend:
    """The C-ABI calls of one row-sharded step (what the stages above drive)."""

    def __init__(self, plan, n_layers, d, leaky):
        from . import _lib
        self.plan, self.L, self.d, self.leaky = plan, n_layers, d, float(leaky)
        self.lib, self._lib = _lib.load_library(), _lib
        self.ws, mask_bytes = plan.scratch(n_layers, d)
        self.mask_bytes = mask_bytes

    def begin_forward(self, u, i, u_out, i_out, need_bwd):
        self.u, self.i, self.u_out, self.i_out = u, i, u_out, i_out
        self.masks = torch.empty(max(self.mask_bytes, 1), dtype=torch.uint8, device=u.device) if need_bwd else None

    def begin_backward(self, gu, gi, d_u, d_i):
        self.gu, self.gi, self.d_u, self.d_i = gu, gi, d_u, d_i

    def fwd_layers(self, a, b):
        from .propagate import _ptr, _stream_ptr
        p = self.plan
        with torch.cuda.device(p.device):
            self._lib.check(self.lib.sagnn_propagate_fwd_layers(
                p.handle, a, b, _ptr(self.u), _ptr(self.i), _ptr(self.u_out), _ptr(self.i_out), self.L, self.d,
                self.leaky, _ptr(self.masks), _ptr(self.ws), self.ws.numel(), _stream_ptr(p.device)))

    def bwd_levels(self, a, b):
        from .propagate import _ptr, _stream_ptr
        p = self.plan
        with torch.cuda.device(p.device):
            self._lib.check(self.lib.sagnn_propagate_bwd_levels(
                p.handle, a, b, _ptr(self.gu), _ptr(self.gi), _ptr(self.d_u), _ptr(self.d_i), self.L, self.d,
                self.leaky, _ptr(self.masks), _ptr(self.ws), self.ws.numel(), _stream_ptr(p.device)))

    def table(self, which, index):
        return self.plan.workspace_table(self.ws, self.L, self.d, which, index)


class RowShardedPropagation:
    """Row-sharded drop-in for ``propagate`` over a process group: every rank passes the FULL
    ``uEmbed [T,U,d]`` / ``iEmbed [T,I,d]`` (replicated parameters) and gets the full layer sums
    back; each rank computes 1/world of the rows of every interval.  Use it when there are fewer
    intervals than GPUs (or combine: interval-shard over groups of ranks, row-shard inside a group
    by passing ``group``).  Results are bitwise those of the single-GPU ``propagate``."""

    def __init__(self, sub_mats, U, I, n_layers=2, leaky=0.5, group=None, device=None, latdim=64):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.U, self.I, self.n_layers, self.leaky = int(U), int(I), int(n_layers), float(leaky)
        self.U_pad, self.I_pad = padded_rows(self.U, self.world), padded_rows(self.I, self.world)
        bu, bi = self.U_pad // self.world, self.I_pad // self.world
        self.row_block = (self.rank * bu, (self.rank + 1) * bu, self.rank * bi, (self.rank + 1) * bi)
        self._backends = {}
        self.plan = self._build_plan(sub_mats, device, latdim)

    # The two hooks below are the only places that touch the CUDA library; there is no CPU path in the product.
    # (tests/test_dist_cpu.py subclasses this class and overrides them with a dense stand-in to drive the
    # exchange schedule over gloo.)
    def _build_plan(self, sub_mats, device, latdim):
        from .propagate import build_plan
        return build_plan(sub_mats, self.U, self.I, device=device, latdim=latdim,
                          row_block=self.row_block, padded_shape=(self.U_pad, self.I_pad))

    def _make_backend(self, d):
        return _CudaRowBackend(self.plan, self.n_layers, d, self.leaky)

    def _backend(self, d):
        if d not in self._backends:
            self._backends[d] = self._make_backend(d)
        return self._backends[d]

    @staticmethod
    def _pad(t, rows, rows_pad):
        t = t.contiguous()
        if rows_pad == rows:
            return t
        out = t.new_zeros((t.shape[0], rows_pad, t.shape[2]))
        out[:, :rows] = t
        return out

    def __call__(self, u_embed, i_embed):
        return _RowShardedFn.apply(u_embed, i_embed, self)
