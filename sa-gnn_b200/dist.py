"""Multi-GPU plumbing for the propagation path: interval sharding + output all-gather.

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


class ShardedPropagation:
    """Interval-sharded drop-in for ``propagate``: every rank passes the FULL parameter tables
    (or just its own slices via ``local_only``) and gets the full ``[T,U,d]`` / ``[T,I,d]`` back."""

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
