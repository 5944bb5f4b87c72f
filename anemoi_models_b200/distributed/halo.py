"""dst-node sharding with a halo exchange of src rows.

Replaces, on the graph path, the reference's Ulysses-style head all-to-all (block.py:366-414: four all-to-alls
moving q, k, v and the E x D edge features, plus one for the output) and `sync_tensor`'s full all-gather
(block.py:203, 259-260).  Each rank owns a contiguous range of dst rows (the same `tensor_split` ranges the
model's shard shapes describe) together with every edge that points into it; edge features never leave the GPU.
The only exchange is the set of src rows a rank's edges reference but another rank owns ("halo"), moved with one
`all_to_all_single` (NCCL over NVLink); backward sends the gradients of those rows back and adds them at the owner.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .collectives import all_gather_rows, reduce_scatter_rows


@dataclass
class HaloPlan:
    rank: int
    world: int
    edge_ids: Tensor           # [E_r] original ids of the edges this rank owns (original order = reference chunk r)
    local_edge_index: Tensor   # [2, E_r] row 0 = index into the compact src buffer, row 1 = dst - dst_lo
    n_needed: int              # rows of the compact src buffer (= sum(recv_counts))
    send_idx: Tensor           # [sum(send_counts)] rows of the OWN src shard to send, grouped by peer
    send_counts: List[int]
    recv_counts: List[int]
    num_dst_local: int
    dst_lo: int
    num_src_local: int
    src_lo: int
    csr: object = field(default=None, repr=False)  # GraphCSR of the local graph, built lazily on CUDA


def _unique_sorted(x: Tensor) -> Tensor:
    return torch.unique(x, sorted=True)


def build_bipartite_halo_plan(edge_index: Tensor, src_bounds: List[int], dst_bounds: List[int], rank: int) -> HaloPlan:
    """Plan from the FULL edge_index, which the GT mappers/processors replicate on every rank
    (reference mapper.py:254, processor.py:326-331), so no communication is needed to build it.

    Bit-exact contract: `edge_ids` equals chunk `rank` of the reference's sort_edges_1hop_chunks
    (khop_edges.py:88-130) for the dst ranges given by `dst_bounds`."""
    P = len(dst_bounds) - 1
    src, dst = edge_index[0].long(), edge_index[1].long()
    src_lo, src_hi = src_bounds[rank], src_bounds[rank + 1]
    send_lists, send_counts = [], []
    needed_mine = None
    mask_mine = None
    for p in range(P):
        mask = (dst >= dst_bounds[p]) & (dst < dst_bounds[p + 1])
        needed = _unique_sorted(src[mask])
        if p == rank:
            needed_mine, mask_mine = needed, mask
        own = needed[(needed >= src_lo) & (needed < src_hi)] - src_lo
        send_lists.append(own)
        send_counts.append(int(own.numel()))
    sb = torch.tensor(src_bounds, device=edge_index.device, dtype=torch.long)
    owner = torch.searchsorted(sb, needed_mine, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=P)[:P].tolist()
    edge_ids = mask_mine.nonzero(as_tuple=False).view(-1)
    local_src = torch.searchsorted(needed_mine, src[edge_ids])
    local_dst = dst[edge_ids] - dst_bounds[rank]
    return HaloPlan(rank=rank, world=P, edge_ids=edge_ids, local_edge_index=torch.stack([local_src, local_dst]).contiguous(),
                    n_needed=int(needed_mine.numel()), send_idx=torch.cat(send_lists) if send_lists else src.new_zeros(0),
                    send_counts=send_counts, recv_counts=[int(c) for c in recv_counts],
                    num_dst_local=dst_bounds[rank + 1] - dst_bounds[rank], dst_lo=dst_bounds[rank],
                    num_src_local=src_hi - src_lo, src_lo=src_lo)


def build_local_halo_plan(edge_index_local: Tensor, src_bounds: List[int], dst_bounds: List[int], group) -> HaloPlan:
    """Plan when each rank only holds ITS edges (global node ids) -- the GraphConv path, where the processor has
    already partitioned the edges with sort_edges_1hop_sharding (reference processor.py:239-246).  The lists of
    needed rows are exchanged once (two small all-to-alls) and the plan is cached by the caller."""
    rank, P = dist.get_rank(group=group), dist.get_world_size(group=group)
    dev = edge_index_local.device
    src, dst = edge_index_local[0].long(), edge_index_local[1].long()
    needed = _unique_sorted(src)
    sb = torch.tensor(src_bounds, device=dev, dtype=torch.long)
    owner = torch.searchsorted(sb, needed, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=P)[:P]
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts, group=group)
    send_counts_l, recv_counts_l = send_counts.tolist(), recv_counts.tolist()
    send_idx = torch.empty(sum(send_counts_l), dtype=torch.long, device=dev)
    dist.all_to_all_single(send_idx, needed.contiguous(), send_counts_l, recv_counts_l, group=group)
    send_idx = send_idx - src_bounds[rank]
    local_src = torch.searchsorted(needed, src)
    local_dst = dst - dst_bounds[rank]
    return HaloPlan(rank=rank, world=P, edge_ids=torch.arange(src.numel(), device=dev),
                    local_edge_index=torch.stack([local_src, local_dst]).contiguous(), n_needed=int(needed.numel()),
                    send_idx=send_idx, send_counts=[int(c) for c in send_counts_l], recv_counts=[int(c) for c in recv_counts_l],
                    num_dst_local=dst_bounds[rank + 1] - dst_bounds[rank], dst_lo=dst_bounds[rank],
                    num_src_local=src_bounds[rank + 1] - src_bounds[rank], src_lo=src_bounds[rank])


class _HaloGather(torch.autograd.Function):
    """fwd: compact buffer of the src rows this rank's edges need ([own rows | halo rows], ascending global id);
    bwd: gradients of those rows travel back to their owners and are summed there (fp32 accumulation)."""

    @staticmethod
    def forward(ctx, x: Tensor, plan: HaloPlan, group) -> Tensor:
        ctx.plan, ctx.group = plan, group
        ctx.rows = x.shape[0]
        send = x.index_select(0, plan.send_idx).contiguous()
        recv = torch.empty((plan.n_needed,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recv, send, plan.recv_counts, plan.send_counts, group=group)
        return recv

    @staticmethod
    def backward(ctx, g: Tensor):
        plan: HaloPlan = ctx.plan
        g = g.contiguous()
        back = torch.empty((sum(plan.send_counts),) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.all_to_all_single(back, g, plan.send_counts, plan.recv_counts, group=ctx.group)
        dx = torch.zeros((ctx.rows,) + tuple(g.shape[1:]), dtype=torch.float32, device=g.device)
        dx.index_add_(0, plan.send_idx, back.float())
        return dx.to(g.dtype), None, None


def halo_gather(x: Tensor, plan: HaloPlan, group) -> Tensor:
    """[n_needed, ...] rows of the (row-sharded) tensor x referenced by this rank's edges."""
    return _HaloGather.apply(x, plan, group)


class _SelectShardedEdges(torch.autograd.Function):
    """Raw edge attributes arrive sharded by `tensor_split` over the ORIGINAL edge order (reference mapper.py:255-256);
    this rank needs the rows of the edges whose dst it owns.  fwd: all-gather the [E, edge_dim] rows (edge_dim ~ 11, so
    this is ~1 % of what the reference's head all-to-all of the projected E x D features moves) and pick own rows;
    bwd: scatter into a zero [E, edge_dim] buffer, sum over ranks, keep own shard -- the dual `sync_tensor` uses."""

    @staticmethod
    def forward(ctx, ea: Tensor, shapes, edge_ids: Tensor, group) -> Tensor:
        ctx.shapes, ctx.group = shapes, group
        ctx.save_for_backward(edge_ids)
        full = all_gather_rows(ea, 0, shapes, group)
        return full.index_select(0, edge_ids)

    @staticmethod
    def backward(ctx, g: Tensor):
        (edge_ids,) = ctx.saved_tensors
        total = sum(int(s[0]) for s in ctx.shapes)
        full = torch.zeros((total,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        full.index_copy_(0, edge_ids, g)
        return reduce_scatter_rows(full, 0, ctx.shapes, ctx.group), None, None, None


def select_sharded_edges(edge_attr_shard: Tensor, shapes_edge, edge_ids: Tensor, group) -> Tensor:
    return _SelectShardedEdges.apply(edge_attr_shard, shapes_edge, edge_ids, group)
