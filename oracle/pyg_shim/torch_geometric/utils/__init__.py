"""Restated torch_geometric.utils entry points used on the reference's graph path.

Call sites in the reference: conv.py:21-22,74,139 (scatter, softmax); khop_edges.py:19-21,43-47,121-126
(bipartite_subgraph, k_hop_subgraph, mask_to_index).
"""
from typing import Optional, Tuple, Union

import torch
from torch import Tensor


def _broadcast(index: Tensor, ref: Tensor, dim: int) -> Tensor:
    shape = [1] * ref.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(ref)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    """PyG utils.scatter: `sum`/`add` -> zeros.scatter_add_; `max` -> zeros.scatter_reduce_(amax, include_self=False)."""
    if index.dim() != 1:
        raise ValueError(f"The `index` argument must be one-dimensional (got {index.dim()} dimensions)")
    dim = src.dim() + dim if dim < 0 else dim
    if dim < 0 or dim >= src.dim():
        raise ValueError(f"The `dim` argument must lay between 0 and {src.dim() - 1} (got {dim})")
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce in ("max", "amax"):
        return src.new_zeros(size).scatter_reduce_(dim, _broadcast(index, src, dim), src, reduce="amax", include_self=False)
    if reduce in ("min", "amin"):
        return src.new_zeros(size).scatter_reduce_(dim, _broadcast(index, src, dim), src, reduce="amin", include_self=False)
    if reduce == "mean":
        count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(dim))).clamp_(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
        shape = [1] * out.dim()
        shape[dim] = -1
        return out / count.view(shape)
    raise ValueError(f"Encountered invalid `reduce` argument '{reduce}'")


def softmax(
    src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None, num_nodes: Optional[int] = None, dim: int = 0
) -> Tensor:
    """PyG utils.softmax (index form): max on src.detach(), exp(src-max), / (segment sum + 1e-16)."""
    if ptr is not None:
        raise NotImplementedError("ptr form is never used by the reference (conv.py:139 passes ptr=None)")
    if index is None:
        raise NotImplementedError
    N = num_nodes if num_nodes is not None else (int(index.max()) + 1 if index.numel() > 0 else 0)
    src_max = scatter(src.detach(), index, dim, dim_size=N, reduce="max")
    out = src - src_max.index_select(dim, index)
    out = out.exp()
    out_sum = scatter(out, index, dim, dim_size=N, reduce="sum") + 1e-16
    out_sum = out_sum.index_select(dim, index)
    return out / out_sum


def mask_to_index(mask: Tensor) -> Tensor:
    return mask.nonzero(as_tuple=False).view(-1)


def index_to_mask(index: Tensor, size: Optional[int] = None) -> Tensor:
    index = index.view(-1)
    size = int(index.max()) + 1 if size is None else size
    mask = index.new_zeros(size, dtype=torch.bool)
    mask[index] = True
    return mask


def bipartite_subgraph(
    subset: Tuple[Tensor, Tensor],
    edge_index: Tensor,
    edge_attr: Optional[Tensor] = None,
    relabel_nodes: bool = False,
    size: Optional[Tuple[int, int]] = None,
    return_edge_mask: bool = False,
):
    """Edges whose src is in subset[0] and dst in subset[1]; order preserving, no relabelling (as used at khop_edges.py:121)."""
    if relabel_nodes:
        raise NotImplementedError
    src_subset, dst_subset = subset
    if size is None:
        size = (int(edge_index[0].max()) + 1, int(edge_index[1].max()) + 1)
    if src_subset.dtype != torch.bool:
        src_subset = index_to_mask(src_subset, size[0])
    if dst_subset.dtype != torch.bool:
        dst_subset = index_to_mask(dst_subset, size[1])
    edge_mask = src_subset[edge_index[0]] & dst_subset[edge_index[1]]
    edge_index = edge_index[:, edge_mask]
    edge_attr = edge_attr[edge_mask] if edge_attr is not None else None
    if return_edge_mask:
        return edge_index, edge_attr, edge_mask
    return edge_index, edge_attr


def k_hop_subgraph(
    node_idx: Union[int, list, Tensor],
    num_hops: int,
    edge_index: Tensor,
    relabel_nodes: bool = False,
    num_nodes: Optional[int] = None,
    flow: str = "source_to_target",
    directed: bool = False,
):
    """PyG k_hop_subgraph restated.  The reference only calls it with num_hops=1, directed=True (khop_edges.py:43)."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    assert flow in ("source_to_target", "target_to_source")
    if flow == "target_to_source":
        row, col = edge_index
    else:
        col, row = edge_index
    node_mask = row.new_empty(num_nodes, dtype=torch.bool)
    edge_mask = row.new_empty(row.size(0), dtype=torch.bool)
    if isinstance(node_idx, int):
        node_idx = torch.tensor([node_idx], device=row.device)
    elif isinstance(node_idx, (list, tuple)):
        node_idx = torch.tensor(node_idx, device=row.device)
    else:
        node_idx = node_idx.to(row.device)
    subsets = [node_idx]
    preserved_edge_mask = None
    for _ in range(num_hops):
        node_mask.fill_(False)
        node_mask[subsets[-1]] = True
        torch.index_select(node_mask, 0, row, out=edge_mask)
        subsets.append(col[edge_mask])
        if preserved_edge_mask is None:
            preserved_edge_mask = edge_mask.clone()
        else:
            preserved_edge_mask |= edge_mask
    subset, inv = torch.cat(subsets).unique(return_inverse=True)
    inv = inv[: node_idx.numel()]
    node_mask.fill_(False)
    node_mask[subset] = True
    if not directed:
        edge_mask = node_mask[row] & node_mask[col]
    else:
        edge_mask = preserved_edge_mask
    edge_index = edge_index[:, edge_mask]
    if relabel_nodes:
        raise NotImplementedError
    return subset, edge_index, inv, edge_mask
