"""dst-node sharding with a halo exchange of src rows.

Replaces, on the graph path, the reference's Ulysses-style head all-to-all (block.py:366-414: four all-to-alls
moving q, k, v and the E x D edge features, plus one for the output) and `sync_tensor`'s full all-gather
(block.py:203, 259-260).  Each rank owns a contiguous range of dst rows (the same `tensor_split` ranges the
model's shard shapes describe) together with every edge that points into it; edge features never leave the GPU.
The only exchange is the set of src rows a rank's edges reference but another rank owns ("halo"), moved with one
`all_to_all_single` (NCCL over NVLink); backward sends the gradients of those rows back and adds them at the owner.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .collectives import all_gather_rows, reduce_scatter_rows


@dataclass
class HaloPlan:
    rank: int
    world: int
    edge_ids: Tensor           # [E_r] original ids of the edges this rank owns (original order = reference chunk r)
    local_edge_index: Tensor   # [2, E_r] row 0 = COMPACT src index, row 1 = dst - dst_lo
    n_own: int                 # src rows of the own shard: compact indices [0, n_own) = global id - src_lo
    n_halo: int                # src rows owned by peers: compact indices [n_own, n_own + n_halo), ascending global id
    send_idx: Tensor           # [sum(send_counts)] rows of the OWN src shard the peers need, grouped by peer
    send_counts: List[int]     # per peer (0 for self)
    recv_counts: List[int]     # per peer (0 for self); sum = n_halo
    num_dst_local: int
    dst_lo: int
    src_lo: int
    halo_ids: Tensor = field(default=None, repr=False)  # [n_halo] global ids of the halo rows

    @property
    def n_src(self) -> int:
        return self.n_own + self.n_halo


def _compact_sources(src: Tensor, src_bounds: List[int], rank: int):
    """compact index of every edge's src: own rows keep (id - src_lo), the others are numbered n_own + position in the
    sorted list of distinct non-owned ids.  Returns (compact src, halo ids sorted, owner rank of each halo id)."""
    lo, hi = src_bounds[rank], src_bounds[rank + 1]
    own = (src >= lo) & (src < hi)
    halo_ids = torch.unique(src[~own], sorted=True)
    compact = torch.where(own, src - lo, (hi - lo) + torch.searchsorted(halo_ids, src))
    sb = torch.tensor(src_bounds, device=src.device, dtype=torch.long)
    owner = torch.searchsorted(sb, halo_ids, right=True) - 1
    return compact, halo_ids, owner


def build_bipartite_halo_plan(edge_index: Tensor, src_bounds: List[int], dst_bounds: List[int], rank: int) -> HaloPlan:
    """Plan from the FULL edge_index, which the GT mappers/processors replicate on every rank
    (reference mapper.py:254, processor.py:326-331), so no communication is needed to build it.

    Bit-exact contract: `edge_ids` equals chunk `rank` of the reference's sort_edges_1hop_chunks
    (khop_edges.py:88-130) for the dst ranges given by `dst_bounds`."""
    P = len(dst_bounds) - 1
    src, dst = edge_index[0].long(), edge_index[1].long()
    src_lo, src_hi = src_bounds[rank], src_bounds[rank + 1]
    send_lists, send_counts = [], []
    mask_mine = None
    for p in range(P):
        mask = (dst >= dst_bounds[p]) & (dst < dst_bounds[p + 1])
        if p == rank:
            mask_mine = mask
            send_lists.append(src.new_zeros(0))
            send_counts.append(0)
            continue
        needed = torch.unique(src[mask], sorted=True)  # what peer p's edges reference ...
        mine = needed[(needed >= src_lo) & (needed < src_hi)] - src_lo  # ... of my rows
        send_lists.append(mine)
        send_counts.append(int(mine.numel()))
    edge_ids = mask_mine.nonzero(as_tuple=False).view(-1)
    compact, halo_ids, owner = _compact_sources(src[edge_ids], src_bounds, rank)
    recv_counts = torch.bincount(owner, minlength=P)[:P].tolist()
    local_dst = dst[edge_ids] - dst_bounds[rank]
    return HaloPlan(rank=rank, world=P, edge_ids=edge_ids, local_edge_index=torch.stack([compact, local_dst]).contiguous(),
                    n_own=src_hi - src_lo, n_halo=int(halo_ids.numel()), send_idx=torch.cat(send_lists),
                    send_counts=send_counts, recv_counts=[int(c) for c in recv_counts],
                    num_dst_local=dst_bounds[rank + 1] - dst_bounds[rank], dst_lo=dst_bounds[rank], src_lo=src_lo,
                    halo_ids=halo_ids)


def build_local_halo_plan(edge_index_local: Tensor, src_bounds: List[int], dst_bounds: List[int], group) -> HaloPlan:
    """Plan when each rank only holds ITS edges (global node ids) -- the GraphConv path, where the processor has
    already partitioned the edges with sort_edges_1hop_sharding (reference processor.py:239-246), and the weak-scaling
    bench.  The lists of needed rows are exchanged once (two small all-to-alls); the caller caches the plan."""
    rank, P = dist.get_rank(group=group), dist.get_world_size(group=group)
    dev = edge_index_local.device
    src, dst = edge_index_local[0].long(), edge_index_local[1].long()
    compact, halo_ids, owner = _compact_sources(src, src_bounds, rank)
    recv_counts = torch.bincount(owner, minlength=P)[:P]
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts, group=group)
    send_counts_l, recv_counts_l = send_counts.tolist(), recv_counts.tolist()
    send_idx = torch.empty(sum(send_counts_l), dtype=torch.long, device=dev)
    dist.all_to_all_single(send_idx, halo_ids.contiguous(), send_counts_l, recv_counts_l, group=group)
    send_idx = send_idx - src_bounds[rank]
    local_dst = dst - dst_bounds[rank]
    return HaloPlan(rank=rank, world=P, edge_ids=torch.arange(src.numel(), device=dev),
                    local_edge_index=torch.stack([compact, local_dst]).contiguous(),
                    n_own=src_bounds[rank + 1] - src_bounds[rank], n_halo=int(halo_ids.numel()), send_idx=send_idx,
                    send_counts=[int(c) for c in send_counts_l], recv_counts=[int(c) for c in recv_counts_l],
                    num_dst_local=dst_bounds[rank + 1] - dst_bounds[rank], dst_lo=dst_bounds[rank], src_lo=src_bounds[rank],
                    halo_ids=halo_ids)


_plan_registry: dict = {}  # (id(edge_index), key) -> (weakref to the tensor, version, plan)


def cached_plan(owner, key, edge_index: Tensor, group, builder):
    """Halo plan of `edge_index` kept ON THE CALLING MODULE (`owner.__dict__`), one entry per key (shard bounds, rank).

    Building a plan and, later, its peer-memory exchange involves collectives, so whether to build must be the same decision on
    every rank.  A process-wide LRU keyed by tensor content (the round-1 design) made that decision locally: had eviction
    order or a content match ever differed between ranks, one rank would have entered the collective alone.  Here:
      * the very same tensor object with an unchanged version -> hit, no device work, no host sync (the steady state: the
        expanded edge_index and the 1-hop partition are memoised, see install.py / khop_edges.py);
      * anything else -> every rank compares its tensor with the stored one and the group agrees (MIN all-reduce) whether
        all of them may keep their plan; otherwise all rebuild.  SPMD ranks walk through the same modules in the same
        order, so they reach this point together."""
    from ..graph import tensor_version

    store = owner.__dict__.setdefault("_b200_halo_plans", {})
    hit = store.get(key)
    if hit is not None and hit[0] is edge_index and hit[1] == tensor_version(edge_index):
        return hit[2]
    # another module already holds the plan of this very tensor (the 16 blocks of a processor share one edge_index): share the
    # plan -- and with it the peer-memory buffers -- instead of building 16 copies.  Identity relations between tensors are the
    # same on every rank (same program), so this is again the same decision everywhere.
    reg = _plan_registry.get((id(edge_index), key))
    if reg is not None and reg[0]() is edge_index and reg[1] == tensor_version(edge_index):
        store[key] = (edge_index, reg[1], reg[2])
        return reg[2]
    same = (hit is not None and hit[0].shape == edge_index.shape and hit[0].device == edge_index.device
            and hit[0].dtype == edge_index.dtype and hit[1] == tensor_version(hit[0]) and bool(torch.equal(hit[0], edge_index)))
    flag = torch.tensor([1 if same else 0], dtype=torch.int64, device=edge_index.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        plan = hit[2]
    else:
        plan = builder()
    store[key] = (edge_index, tensor_version(edge_index), plan)
    for k_ in [k_ for k_, v_ in _plan_registry.items() if v_[0]() is None]:
        del _plan_registry[k_]
    _plan_registry[(id(edge_index), key)] = (weakref.ref(edge_index), tensor_version(edge_index), plan)
    return plan


def aligned_bounds_from_ranges(lo: Sequence[int], hi: Sequence[int], n_src: int) -> List[int]:
    """Contiguous src ownership that follows dst ownership: cut between rank r and r+1 in the middle of the zone both
    reference.  lo[r] / hi[r] = smallest / largest src row referenced by the edges into rank r's dst shard (lo > hi: none).

    The reference shards src and dst rows by equal counts (`get_shard_shapes`, distributed/shapes.py:19-29), but passes
    the shard shapes explicitly (block.py:479-550 `shapes`), and the encoder's src tensor is replicated before it is
    sharded (mapper.py:110), so its ownership is the caller's choice.  On latitude-ordered grids of different point
    density (uniform Fibonacci vs octahedral) equal counts do not align in latitude and 9-25 % of a shard becomes halo
    at 4-8 ranks; aligned bounds leave only the cut-off zone (~0.2 %).  Returns P+1 non-decreasing bounds, [0] = 0, [P] = n_src."""
    P = len(lo)
    bounds = [0]
    prev_hi = -1  # largest row referenced by the ranks before the cut
    for r in range(P - 1):
        if lo[r] <= hi[r]:
            prev_hi = max(prev_hi, int(hi[r]))
        nxt = next((int(lo[t]) for t in range(r + 1, P) if lo[t] <= hi[t]), n_src)  # first row referenced after the cut
        cut = (prev_hi + 1 + nxt) // 2  # middle of the zone both sides reference (or of the gap between them)
        bounds.append(min(max(cut, bounds[-1]), n_src))
    bounds.append(n_src)
    return bounds


def aligned_src_bounds(edge_index_local: Tensor, n_src: int, group) -> List[int]:
    """`aligned_bounds_from_ranges` for the edges each rank holds (global ids; rank r holds the edges into ITS dst shard)."""
    P = dist.get_world_size(group=group)
    src = edge_index_local[0]
    mine = torch.stack([src.min(), src.max()]).long() if src.numel() else torch.tensor([1, 0], device=src.device)
    allr = [torch.empty_like(mine) for _ in range(P)]
    dist.all_gather(allr, mine, group=group)
    lohi = torch.stack(allr).cpu().tolist()
    return aligned_bounds_from_ranges([a for a, _ in lohi], [b for _, b in lohi], int(n_src))


def exchange_rows(x: Tensor, plan: HaloPlan, group) -> Tensor:
    """[n_halo, ...] rows of peers' shards this rank's edges reference (one all-to-all; no autograd)."""
    send = x.index_select(0, plan.send_idx).contiguous()
    recv = torch.empty((plan.n_halo,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, send, plan.recv_counts, plan.send_counts, group=group)
    return recv


def return_rows(g_halo: Tensor, plan: HaloPlan, group, into: Tensor) -> Tensor:
    """Send the gradients of halo rows back to their owners and add them into `into` ([n_own, ...], in place).
    Peers are added one after the other; the ids inside one peer's list are distinct, so the sum order is fixed."""
    back = torch.empty((sum(plan.send_counts),) + tuple(g_halo.shape[1:]), dtype=g_halo.dtype, device=g_halo.device)
    dist.all_to_all_single(back, g_halo.contiguous(), plan.send_counts, plan.recv_counts, group=group)
    off = 0
    for cnt in plan.send_counts:
        if cnt:
            into.index_add_(0, plan.send_idx[off:off + cnt], back[off:off + cnt].to(into.dtype))
        off += cnt
    return into


class _HaloExchange(torch.autograd.Function):
    """Differentiable exchange_rows (used for the small node tensors of the GraphConv path and by the CPU tests);
    the GT conv uses the fused `ops.gt_conv_sharded`, which adds the returned gradients in place."""

    @staticmethod
    def forward(ctx, x: Tensor, plan: HaloPlan, group) -> Tensor:
        ctx.plan, ctx.group, ctx.shape = plan, group, tuple(x.shape)
        return exchange_rows(x, plan, group)

    @staticmethod
    def backward(ctx, g: Tensor):
        dx = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        return_rows(g.float(), ctx.plan, ctx.group, dx)
        return dx.to(g.dtype), None, None


def halo_exchange(x: Tensor, plan: HaloPlan, group) -> Tensor:
    return _HaloExchange.apply(x, plan, group)


def halo_gather(x: Tensor, plan: HaloPlan, group) -> Tensor:
    """[n_own + n_halo, ...] = own rows followed by the halo rows (the compact src space of plan.local_edge_index)."""
    return torch.cat([x, halo_exchange(x, plan, group)], dim=0)


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


def select_sharded_edges(edge_attr_shard: Tensor, shapes_edge, edge_ids: Tensor, group, plan: Optional[HaloPlan] = None) -> Tensor:
    """Own rows of the all-gathered raw edge attributes.  A processor hands the SAME `edge_attr` tensor to every one of its
    blocks (reference processor.py:335-341): with `plan` given the gather + select (and, in backward, the reduce-scatter) runs
    once per forward and the result is shared by the blocks that follow (keyed by tensor identity and version; autograd sums
    their gradients into the one gather)."""
    if plan is None or (edge_attr_shard.requires_grad and edge_attr_shard.is_leaf):
        # a leaf that requires grad (a Parameter handed in directly) is the same object in every training step: a result cached
        # from the previous step would hang off that step's (freed) autograd graph
        return _SelectShardedEdges.apply(edge_attr_shard, shapes_edge, edge_ids, group)
    from ..graph import tensor_version

    hit = plan.__dict__.get("_ea_select")
    if (hit is not None and hit[0]() is edge_attr_shard and hit[1] == tensor_version(edge_attr_shard)
            and hit[3] == torch.is_grad_enabled()):
        return hit[2]
    out = _SelectShardedEdges.apply(edge_attr_shard, shapes_edge, edge_ids, group)
    import weakref

    plan.__dict__["_ea_select"] = (weakref.ref(edge_attr_shard), tensor_version(edge_attr_shard), out, torch.is_grad_enabled())
    return out
