"""Shard / gather / sync / reduce with autograd duals (reference distributed/graph.py:20-298, primitives.py:21-143).

Same call signatures and forward/backward pairing as the reference:
    shard_tensor        fwd keep own slice            | bwd all-gather (or zero-fill when gather_in_backward=False)
    gather_tensor       fwd all-gather                | bwd keep own slice
    sync_tensor         fwd all-gather                | bwd sum over ranks, keep own slice
    reduce_tensor       fwd sum over ranks            | bwd identity
    reduce_shard_tensor fwd sum over ranks, own slice | bwd all-gather
All are the identity when the group is falsy or has one rank.  Uneven shards are exchanged through one padded,
fixed-size collective (`all_gather_into_tensor` / `reduce_scatter_tensor` on NCCL) instead of the reference's
list-form collectives, so the same code runs over NCCL on NVLink and over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def _active(group) -> bool:
    return bool(group) and dist.get_world_size(group=group) > 1


def _sizes(shapes, dim: int) -> List[int]:
    return [int(s[dim]) for s in shapes]


def split_rows(x: Tensor, dim: int, shapes, group) -> Tensor:
    """Own slice of x along dim (sections given by `shapes`)."""
    sizes = _sizes(shapes, dim)
    rank = dist.get_rank(group=group)
    start = sum(sizes[:rank])
    return x.narrow(dim, start, sizes[rank]).contiguous()


def all_gather_rows(x: Tensor, dim: int, shapes, group) -> Tensor:
    """Concatenation along dim of every rank's (possibly unevenly sized) shard."""
    P = dist.get_world_size(group=group)
    sizes = _sizes(shapes, dim)
    xm = x.movedim(dim, 0).contiguous()
    if xm.shape[0] != sizes[dist.get_rank(group=group)]:
        raise ValueError(f"local shard has {xm.shape[0]} rows along dim {dim}, shapes say {sizes[dist.get_rank(group=group)]}")
    nmax = max(sizes)
    if all(s == nmax for s in sizes):
        buf = torch.empty((P * nmax,) + tuple(xm.shape[1:]), dtype=xm.dtype, device=xm.device)
        dist.all_gather_into_tensor(buf, xm, group=group)
        out = buf
    else:
        pad = torch.zeros((nmax,) + tuple(xm.shape[1:]), dtype=xm.dtype, device=xm.device)
        pad[: xm.shape[0]] = xm
        buf = torch.empty((P * nmax,) + tuple(xm.shape[1:]), dtype=xm.dtype, device=xm.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        out = torch.cat([buf[p * nmax: p * nmax + sizes[p]] for p in range(P)], dim=0)
    return out.movedim(0, dim).contiguous()


def all_reduce_sum(x: Tensor, group, use_fp32: bool = True) -> Tensor:
    """Sum over ranks; like the reference (primitives.py:135-139) the reduction runs in fp32 by default."""
    if use_fp32 and x.dtype != torch.float32:
        xf = x.float()
        dist.all_reduce(xf, group=group)
        return xf.to(x.dtype)
    xc = x.contiguous().clone()
    dist.all_reduce(xc, group=group)
    return xc


def reduce_scatter_rows(x: Tensor, dim: int, shapes, group, use_fp32: bool = True) -> Tensor:
    """Own slice of the sum over ranks of x (full size on every rank)."""
    P = dist.get_world_size(group=group)
    sizes = _sizes(shapes, dim)
    backend = dist.get_backend(group)
    if backend == "nccl":
        rank = dist.get_rank(group=group)
        xm = x.movedim(dim, 0)
        work_dtype = torch.float32 if use_fp32 else x.dtype
        nmax = max(sizes)
        buf = torch.zeros((P, nmax) + tuple(xm.shape[1:]), dtype=work_dtype, device=x.device)
        start = 0
        for p in range(P):
            buf[p, : sizes[p]] = xm[start: start + sizes[p]]
            start += sizes[p]
        out = torch.empty((nmax,) + tuple(xm.shape[1:]), dtype=work_dtype, device=x.device)
        dist.reduce_scatter_tensor(out, buf.view((P * nmax,) + tuple(xm.shape[1:])), group=group)
        return out[: sizes[rank]].to(x.dtype).movedim(0, dim).contiguous()
    return split_rows(all_reduce_sum(x, group, use_fp32), dim, shapes, group)


class _Shard(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, shapes, gather_in_backward, group):
        ctx.dim, ctx.shapes, ctx.gib, ctx.group = dim, shapes, gather_in_backward, group
        return split_rows(x, dim, shapes, group)

    @staticmethod
    def backward(ctx, g):
        if ctx.gib:
            return all_gather_rows(g, ctx.dim, ctx.shapes, ctx.group), None, None, None, None
        # own rows only; the other ranks' rows are never read downstream (the reference leaves them uninitialised)
        sizes = _sizes(ctx.shapes, ctx.dim)
        rank = dist.get_rank(group=ctx.group)
        shape = list(g.shape)
        shape[ctx.dim] = sum(sizes)
        full = g.new_zeros(shape)
        full.narrow(ctx.dim, sum(sizes[:rank]), sizes[rank]).copy_(g)
        return full, None, None, None, None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, shapes, group):
        ctx.dim, ctx.shapes, ctx.group = dim, shapes, group
        return all_gather_rows(x, dim, shapes, group)

    @staticmethod
    def backward(ctx, g):
        return split_rows(g, ctx.dim, ctx.shapes, ctx.group), None, None, None


class _Sync(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, shapes, group):
        ctx.dim, ctx.shapes, ctx.group = dim, shapes, group
        return all_gather_rows(x, dim, shapes, group)

    @staticmethod
    def backward(ctx, g):
        return reduce_scatter_rows(g, ctx.dim, ctx.shapes, ctx.group), None, None, None


class _Reduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        return all_reduce_sum(x, group)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ReduceShard(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, shapes, group):
        ctx.dim, ctx.shapes, ctx.group = dim, shapes, group
        return reduce_scatter_rows(x, dim, shapes, group)

    @staticmethod
    def backward(ctx, g):
        return all_gather_rows(g, ctx.dim, ctx.shapes, ctx.group), None, None, None


def shard_tensor(input_: Tensor, dim: int, shapes: tuple, mgroup, gather_in_backward: bool = True) -> Tensor:
    """reference graph.py:20-44"""
    return _Shard.apply(input_, dim, shapes, gather_in_backward, mgroup) if _active(mgroup) else input_


def gather_tensor(input_: Tensor, dim: int, shapes: tuple, mgroup) -> Tensor:
    """reference graph.py:47-68"""
    return _Gather.apply(input_, dim, shapes, mgroup) if _active(mgroup) else input_


def reduce_tensor(input_: Tensor, mgroup) -> Tensor:
    """reference graph.py:71-88"""
    return _Reduce.apply(input_, mgroup) if _active(mgroup) else input_


def sync_tensor(input_: Tensor, dim: int, shapes: tuple, mgroup) -> Tensor:
    """reference graph.py:91-112"""
    return _Sync.apply(input_, dim, shapes, mgroup) if _active(mgroup) else input_


def reduce_shard_tensor(input_: Tensor, dim: int, shapes: tuple, mgroup) -> Tensor:
    """reference graph.py:115-137"""
    return _ReduceShard.apply(input_, dim, shapes, mgroup) if _active(mgroup) else input_
