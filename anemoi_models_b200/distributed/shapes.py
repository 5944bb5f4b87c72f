"""Shard-shape bookkeeping (reference distributed/shapes.py:19-29); integer, bit-exact."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def tensor_split_sizes(n: int, parts: int) -> List[int]:
    """Section lengths of torch.tensor_split(x, parts): the first n % parts sections get one extra row."""
    base, rem = divmod(int(n), int(parts))
    return [base + 1 if r < rem else base for r in range(parts)]


def group_size(group) -> int:
    return 1 if not group else dist.get_world_size(group=group)


def get_shape_shards(tensor: Tensor, dim: int, model_comm_group=None) -> list:
    """Shapes of the shards `torch.tensor_split(tensor, P, dim)` would produce (reference shapes.py:19-24)."""
    assert dim < tensor.dim(), f"Error, tensor dimension is {tensor.dim()} which cannot be split along {dim}"
    out = []
    for s in tensor_split_sizes(tensor.shape[dim], group_size(model_comm_group)):
        shape = list(tensor.shape)
        shape[dim] = s
        out.append(shape)
    return out


def change_channels_in_shape(shape_list: list, channels: int) -> list:
    """reference shapes.py:27-29"""
    return [x[:-1] + [channels] for x in shape_list] if shape_list else []


def bounds_from_shapes(shapes: list, dim: int = 0) -> List[int]:
    """[0, n0, n0+n1, ...] row offsets of the shards described by a shape list."""
    out = [0]
    for s in shapes:
        out.append(out[-1] + int(s[dim]))
    return out
