"""Head <-> sequence re-sharding (Ulysses-style all-to-all): API mirror of the reference's distributed/transformer.py.

The graph path of this package does NOT use these on its step path -- `GraphTransformer*Block.forward` shards by dst rows and
exchanges a halo instead (distributed/halo.py), which keeps the E x D edge features on their GPU.  They exist so that code written
against the reference (`shard_heads` transformer.py:85-108, `shard_sequence` :111-133, the blocks' `shard_qkve_heads` /
`shard_output_seq` block.py:366-414) keeps working with a `model_comm_group`.

Layout contract (same as the reference): tensors are `(batch..., heads, sequence, channels)`.
  shard_heads   : every rank holds ALL heads of ITS sequence shard  ->  ITS heads of the WHOLE sequence
  shard_sequence: the inverse; each is the other's backward.
Heads are divided like `torch.tensor_split(heads, P)` (uneven counts allowed); sequence shards follow `shapes`
(`shapes[r][0]` rows on rank r).  Differences from the reference, both only visible where the reference fails:
its sequence all-to-all assumes every rank holds as many heads as the caller and splits the sequence evenly whatever `shapes`
says (transformer.py:59-82) -- here both sides use the real counts.  One `all_to_all_single` per call (NCCL or gloo).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from .shapes import tensor_split_sizes


def _active(group) -> bool:
    return bool(group) and dist.is_available() and dist.is_initialized() and dist.get_world_size(group=group) > 1


def get_memory_format(t: Tensor) -> torch.memory_format:
    """channels_last for 4-D channels-last tensors, contiguous otherwise (reference distributed/utils.py:15-32)."""
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous():
        return torch.channels_last
    return torch.contiguous_format


def _seq_sizes(shapes: Sequence, world: int) -> List[int]:
    sizes = [int(s[0]) for s in shapes]
    if len(sizes) != world:
        raise ValueError(f"shapes describes {len(sizes)} shards but the group has {world} ranks")
    return sizes


def _all_head_counts(h_local: int, group, device) -> List[int]:
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(dist.get_world_size(group=group))]
    dist.all_gather(counts, torch.tensor([h_local], dtype=torch.int64, device=device), group=group)
    return [int(c) for c in counts]


def _heads_to_all(x: Tensor, seq: List[int], heads: List[int], group) -> Tensor:
    """(..., H, n_r, C) on rank r  ->  (..., heads[r], sum(seq), C): my heads of every rank's sequence shard."""
    rank = dist.get_rank(group=group)
    fmt = get_memory_format(x)
    lead = tuple(x.shape[:-3])
    H, n_r, C = x.shape[-3:]
    if sum(heads) != H or n_r != seq[rank]:
        raise ValueError(f"shard_heads: tensor has {H} heads / {n_r} rows, expected {sum(heads)} / {seq[rank]}")
    nlead = 1
    for d in lead:
        nlead *= int(d)
    send = x.movedim(-3, 0).contiguous().view(-1)  # heads outermost: peer p's part is one contiguous slice
    send_counts = [h * nlead * n_r * C for h in heads]
    recv_counts = [heads[rank] * nlead * s * C for s in seq]
    recv = torch.empty(sum(recv_counts), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, send, recv_counts, send_counts, group=group)
    parts = [p.view(heads[rank], *lead, s, C) for p, s in zip(torch.split(recv, recv_counts), seq)]
    return torch.cat(parts, dim=-2).movedim(0, -3).contiguous(memory_format=fmt)


def _seq_to_all(x: Tensor, seq: List[int], heads: List[int], group) -> Tensor:
    """(..., heads[r], N, C) on rank r  ->  (..., sum(heads), seq[r], C): all heads of my sequence shard."""
    rank = dist.get_rank(group=group)
    fmt = get_memory_format(x)
    lead = tuple(x.shape[:-3])
    h_r, N, C = x.shape[-3:]
    if h_r != heads[rank] or N != sum(seq):
        raise ValueError(f"shard_sequence: tensor has {h_r} heads / {N} rows, expected {heads[rank]} / {sum(seq)}")
    nlead = 1
    for d in lead:
        nlead *= int(d)
    send = torch.cat([p.contiguous().view(-1) for p in torch.split(x, seq, dim=-2)])  # peer p gets its rows of my heads
    send_counts = [nlead * h_r * s * C for s in seq]
    recv_counts = [nlead * h * seq[rank] * C for h in heads]
    recv = torch.empty(sum(recv_counts), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, send, recv_counts, send_counts, group=group)
    parts = [p.view(*lead, h, seq[rank], C) for p, h in zip(torch.split(recv, recv_counts), heads)]
    return torch.cat(parts, dim=-3).contiguous(memory_format=fmt)


class _ShardHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, shapes, group):
        ctx.group = group if _active(group) else None
        if ctx.group is None:
            return x
        world = dist.get_world_size(group=group)
        ctx.seq, ctx.heads = _seq_sizes(shapes, world), tensor_split_sizes(int(x.shape[-3]), world)
        return _heads_to_all(x, ctx.seq, ctx.heads, group)

    @staticmethod
    def backward(ctx, g: Tensor):
        if ctx.group is None:
            return g, None, None
        return _seq_to_all(g.contiguous(), ctx.seq, ctx.heads, ctx.group), None, None


class _ShardSequence(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, shapes, group):
        ctx.group = group if _active(group) else None
        if ctx.group is None:
            return x
        world = dist.get_world_size(group=group)
        ctx.seq, ctx.heads = _seq_sizes(shapes, world), _all_head_counts(int(x.shape[-3]), group, x.device)
        return _seq_to_all(x, ctx.seq, ctx.heads, group)

    @staticmethod
    def backward(ctx, g: Tensor):
        if ctx.group is None:
            return g, None, None
        return _heads_to_all(g.contiguous(), ctx.seq, ctx.heads, ctx.group), None, None


def shard_heads(input_: Tensor, shapes: list, mgroup: Optional[object]) -> Tensor:
    """(batch..., heads, seq_shard, channels) -> (batch..., head_shard, seq, channels); identity without a group."""
    return _ShardHeads.apply(input_, shapes, mgroup)


def shard_sequence(input_: Tensor, shapes: list, mgroup: Optional[object]) -> Tensor:
    """(batch..., head_shard, seq, channels) -> (batch..., heads, seq_shard, channels); identity without a group."""
    return _ShardSequence.apply(input_, shapes, mgroup)
