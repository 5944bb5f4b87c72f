"""TEST INFRASTRUCTURE ONLY -- integer restatement of the reference's shard / edge-partition helpers.

Follows distributed/shapes.py:19-29 and distributed/khop_edges.py:24-130 of the reference
(/root/reference/src/anemoi/models/).  Everything here is bit-exact integer work in numpy.
"""
from __future__ import annotations

from typing import List, Tuple, Union

import numpy as np


def tensor_split_sizes(n: int, parts: int) -> List[int]:
    """torch.tensor_split(x, parts) section lengths: the first n % parts sections get one extra row."""
    base, rem = divmod(n, parts)
    return [base + 1 if r < rem else base for r in range(parts)]


def shape_shards(shape: Tuple[int, ...], dim: int, parts: int) -> List[List[int]]:
    """get_shape_shards (shapes.py:19-24): shapes of torch.tensor_split(tensor, parts, dim)."""
    out = []
    for s in tensor_split_sizes(shape[dim], parts):
        sh = list(shape)
        sh[dim] = s
        out.append(sh)
    return out


def change_channels(shape_list: List[List[int]], channels: int) -> List[List[int]]:
    """change_channels_in_shape (shapes.py:27-29)."""
    return [x[:-1] + [channels] for x in shape_list] if shape_list else []


def edges_1hop_chunks(num_nodes: Union[int, Tuple[int, int]], edge_index: np.ndarray, num_chunks: int) -> List[np.ndarray]:
    """sort_edges_1hop_chunks (khop_edges.py:88-130), returned as the list of ORIGINAL edge ids per chunk.

    dst range = arange(Nd).tensor_split(num_chunks); chunk c keeps, in original edge order, the edges whose
    dst lies in range c (int num_nodes -> PyG k_hop_subgraph(directed=True, 1 hop) mask = node_mask[dst];
    tuple -> PyG bipartite_subgraph with every src allowed).  The reference then takes
    edge_index[:, ids] / edge_attr[ids]."""
    nd = num_nodes if isinstance(num_nodes, int) else num_nodes[1]
    ei = np.asarray(edge_index)
    dst = ei[1].astype(np.int64)
    sizes = tensor_split_sizes(nd, num_chunks)
    bounds = np.concatenate([[0], np.cumsum(sizes)])
    out = []
    for c in range(num_chunks):
        mask = (dst >= bounds[c]) & (dst < bounds[c + 1])
        out.append(np.nonzero(mask)[0].astype(np.int64))
    return out


def edges_1hop_sharding(num_nodes, edge_index: np.ndarray, parts: int):
    """sort_edges_1hop_sharding (khop_edges.py:50-85) without a process group: returns
    (concatenated original edge ids, per-rank edge counts)."""
    chunks = edges_1hop_chunks(num_nodes, edge_index, parts)
    return np.concatenate(chunks) if chunks else np.zeros(0, np.int64), [int(c.shape[0]) for c in chunks]


def expand_edges(edge_index: np.ndarray, src_size: int, dst_size: int, batch_size: int) -> np.ndarray:
    """GraphEdgeMixin._expand_edges (layers/mapper.py:150-171): cat([edge_index + i*[[Ns],[Nd]] for i in range(B)], 1)."""
    inc = np.array([[src_size], [dst_size]], dtype=np.int64)
    return np.concatenate([np.asarray(edge_index, dtype=np.int64) + i * inc for i in range(batch_size)], axis=1)


def halo_plan(edge_index: np.ndarray, num_src: int, num_dst: int, parts: int):
    """Plan of the dst-sharded halo exchange that replaces the reference's head all-to-all
    (block.py:366-414) / sync_tensor (block.py:203): rank r owns dst rows tensor_split(arange(Nd), P)[r]
    and src rows tensor_split(arange(Ns), P)[r]; it needs every src row referenced by an edge whose dst it owns.
    Returns per rank: dict(edge_ids, needed_src (sorted unique), recv_from[p] = sorted src ids owned by p != r)."""
    ei = np.asarray(edge_index).astype(np.int64)
    chunks = edges_1hop_chunks((num_src, num_dst), ei, parts)
    sb = np.concatenate([[0], np.cumsum(tensor_split_sizes(num_src, parts))])
    plans = []
    for r in range(parts):
        ids = chunks[r]
        needed = np.unique(ei[0, ids])
        owner = np.searchsorted(sb, needed, side="right") - 1
        recv = {p: needed[owner == p] for p in range(parts)}
        plans.append({"edge_ids": ids, "needed_src": needed, "recv_from": recv})
    return plans
