"""1-hop edge partition by dst chunk (reference distributed/khop_edges.py:50-130), on the GPU through
`ab2_edge_chunks`.  Bit-exact: chunk c holds, in original edge order, the edges whose dst lies in
`arange(Nd).tensor_split(num_chunks)[c]`."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import Tensor

from .. import _lib
from .shapes import tensor_split_sizes


def edge_chunk_order(num_dst: int, edge_index: Tensor, num_chunks: int, bounds: Optional[List[int]] = None) -> Tuple[Tensor, List[int]]:
    """(order[int64, E'] original edge ids grouped by chunk, counts per chunk).  One host sync for the counts."""
    if not edge_index.is_cuda:
        raise RuntimeError("anemoi_models_b200 runs on CUDA tensors only (no CPU fallback): edge_index is on the CPU")
    L = _lib.lib()
    ei = edge_index.to(torch.int64).contiguous()
    E = ei.shape[1]
    if bounds is None:
        bounds = [0]
        for s in tensor_split_sizes(num_dst, num_chunks):
            bounds.append(bounds[-1] + s)
    import ctypes as C

    b = (C.c_int64 * (num_chunks + 1))(*bounds)
    order = torch.empty(E, dtype=torch.int64, device=ei.device)
    counts = torch.zeros(num_chunks, dtype=torch.int64, device=ei.device)
    ws_bytes = L.ab2_edge_chunks_workspace_bytes(E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ei.device)
    with torch.cuda.device(ei.device):
        _lib.check(L.ab2_edge_chunks(_lib.ptr(ei), E, C.addressof(b), num_chunks, _lib.ptr(order), _lib.ptr(counts),
                                     _lib.ptr(ws), ws_bytes, _lib.current_stream(ei.device)))
    cnt = counts.tolist()
    return order[: sum(cnt)], cnt


def get_k_hop_edges(nodes: Tensor, edge_attr: Tensor, edge_index: Tensor, num_hops: int = 1) -> Tuple[Tensor, Tensor]:
    """reference khop_edges.py:24-47: (edge_attr, edge_index) of the edges that reach `nodes` within `num_hops` hops
    (flow source -> target, directed), in original edge order.  One hop = the edges whose dst is in `nodes`; every further hop
    adds the edges whose dst is a src reached so far (PyG `k_hop_subgraph(directed=True)`: the union of the per-hop masks)."""
    if num_hops < 1:
        raise ValueError("num_hops must be >= 1")
    src, dst = edge_index[0], edge_index[1]
    n = int(max(int(edge_index.max()) + 1 if edge_index.numel() else 0, int(nodes.max()) + 1 if nodes.numel() else 0))
    frontier = nodes.to(edge_index.device).view(-1)
    keep = torch.zeros(edge_index.shape[1], dtype=torch.bool, device=edge_index.device)
    for _ in range(num_hops):
        node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
        node_mask[frontier] = True
        hop = node_mask[dst]
        keep |= hop
        frontier = src[hop]
    ids = keep.nonzero().view(-1)
    return edge_attr[ids], edge_index[:, ids]


def sort_edges_1hop_chunks(num_nodes: Union[int, Tuple[int, int]], edge_attr: Tensor, edge_index: Tensor,
                           num_chunks: int) -> Tuple[List[Tensor], List[Tensor]]:
    """reference khop_edges.py:88-130: (list of edge_attr chunks, list of edge_index chunks)."""
    nd = num_nodes if isinstance(num_nodes, int) else num_nodes[1]
    order, cnt = edge_chunk_order(nd, edge_index, num_chunks)
    ids = torch.split(order, cnt)
    return [edge_attr[i] for i in ids], [edge_index[:, i] for i in ids]


def sort_edges_1hop_sharding(num_nodes: Union[int, Tuple[int, int]], edge_attr: Tensor, edge_index: Tensor, mgroup=None):
    """reference khop_edges.py:50-85: (edge_attr sorted, edge_index sorted, edge_attr shapes, edge_index shapes)."""
    if mgroup:
        num_chunks = dist.get_world_size(group=mgroup)
        ea_list, ei_list = sort_edges_1hop_chunks(num_nodes, edge_attr, edge_index, num_chunks)
        return (torch.cat(ea_list, dim=0), torch.cat(ei_list, dim=1), [x.shape for x in ea_list], [x.shape for x in ei_list])
    return edge_attr, edge_index, [], []
