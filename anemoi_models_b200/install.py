"""Plug the B200 path into an installed `anemoi.models` (the reference package).

The reference binds the conv and block classes by name at import time (block.py:32-33, mapper.py:30-31,
chunk.py:23-25) and builds mappers/processors through hydra `_target_`s (models/encoder_processor_decoder.py:69-98).
`install()` rebinds those names to the drop-in classes of this package BEFORE the model is constructed, so
`AnemoiModelEncProcDec` picks the new path up unchanged; `uninstall()` restores the originals.
"""
from __future__ import annotations

import importlib
from typing import Dict, Tuple

_saved: Dict[Tuple[str, str], object] = {}

_CONV_NAMES = ("GraphTransformerConv", "GraphConv")
_BLOCK_NAMES = ("GraphTransformerMapperBlock", "GraphTransformerProcessorBlock", "GraphConvMapperBlock",
                "GraphConvProcessorBlock", "GraphTransformerBaseBlock", "GraphConvBaseBlock")
_DIST_NAMES = ("sort_edges_1hop_chunks", "sort_edges_1hop_sharding")


def _rebind(module_name: str, attr: str, new) -> None:
    try:
        mod = importlib.import_module(module_name)
    except ImportError:
        return
    if hasattr(mod, attr):
        _saved.setdefault((module_name, attr), getattr(mod, attr))
        setattr(mod, attr, new)


def _cached_expand_edges(self, edge_index, edge_inc, batch_size: int):
    """GraphEdgeMixin._expand_edges (reference layers/mapper.py:150-171) memoised per (edge_index buffer, batch size).

    The reference rebuilds `cat([edge_index + i * edge_inc ...])` on every forward although both operands are constant
    buffers; returning the SAME tensor object lets the graph-plan cache hit by identity (no rebuild, no device compare,
    no host sync on the step path).  An in-place edit of the buffer (its `_version` changes) rebuilds."""
    import torch

    from .graph import tensor_version

    cache = self.__dict__.setdefault("_b200_expanded", {})
    key = (edge_index.data_ptr(), tensor_version(edge_index), tuple(edge_index.shape), str(edge_index.device), int(batch_size))
    hit = cache.get(key)
    if hit is not None and hit[0] is edge_index and tensor_version(hit[1]) == hit[2]:
        return hit[1]
    # built OUTSIDE inference mode even when the caller is inside it (Lightning validate / predict): the cached tensor is
    # then an ordinary tensor that later training forwards may use, and its version counter exists
    with torch.inference_mode(False):
        out = torch.cat([edge_index + i * edge_inc for i in range(batch_size)], dim=1)
    cache.clear()
    cache[key] = (edge_index, out, tensor_version(out))
    return out


def install(blocks: bool = True, edge_partition: bool = True, cache_expanded_edges: bool = True) -> None:
    """Rebind the reference's symbols.  `blocks=False` swaps only the two conv classes (single-GPU use);
    `blocks=True` also swaps the block classes, which is what enables dst-row sharding with a halo exchange when a
    model_comm_group is passed.  `edge_partition` routes sort_edges_1hop_* through the GPU partition kernel;
    `cache_expanded_edges` memoises GraphEdgeMixin._expand_edges (constant buffers) so the graph plan is found by identity."""
    from . import distributed as b2dist
    from .layers import block as b2block
    from .layers import conv as b2conv

    for name in _CONV_NAMES:
        new = getattr(b2conv, name)
        _rebind("anemoi.models.layers.conv", name, new)
        _rebind("anemoi.models.layers.block", name, new)
    if blocks:
        for name in _BLOCK_NAMES:
            new = getattr(b2block, name)
            for mod in ("anemoi.models.layers.block", "anemoi.models.layers.mapper", "anemoi.models.layers.chunk"):
                _rebind(mod, name, new)
        # the reference's tests check `isinstance(block.node_mlp, MLP)` against the name they import from layers/mlp.py
        # (tests/layers/block/test_block_graphconv.py:50, 86): same constructor, same state_dict keys
        from .layers import mlp as b2mlp

        for mod in ("anemoi.models.layers.mlp", "anemoi.models.layers.block", "anemoi.models.layers.mapper", "anemoi.models.layers.chunk",
                    "anemoi.models.layers.processor", "anemoi.models.layers.conv"):
            _rebind(mod, "MLP", b2mlp.MLP)
    if cache_expanded_edges:
        try:
            mapper_mod = importlib.import_module("anemoi.models.layers.mapper")
            mixin = getattr(mapper_mod, "GraphEdgeMixin", None)
        except ImportError:
            mixin = None
        if mixin is not None and hasattr(mixin, "_expand_edges"):
            _saved.setdefault(("anemoi.models.layers.mapper", "GraphEdgeMixin._expand_edges"), mixin._expand_edges)
            mixin._expand_edges = _cached_expand_edges
    if edge_partition:
        for name in _DIST_NAMES:
            new = getattr(b2dist, name)
            for mod in ("anemoi.models.distributed.khop_edges", "anemoi.models.layers.block",
                        "anemoi.models.layers.mapper", "anemoi.models.layers.processor"):
                _rebind(mod, name, new)


def uninstall() -> None:
    for (module_name, attr), old in list(_saved.items()):
        mod = importlib.import_module(module_name)
        if "." in attr:  # a method of a class of the module
            cls_name, meth = attr.split(".")
            setattr(getattr(mod, cls_name), meth, old)
        else:
            setattr(mod, attr, old)
    _saved.clear()
