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


def install(blocks: bool = True, edge_partition: bool = True) -> None:
    """Rebind the reference's symbols.  `blocks=False` swaps only the two conv classes (single-GPU use);
    `blocks=True` also swaps the block classes, which is what enables dst-row sharding with a halo exchange when a
    model_comm_group is passed.  `edge_partition` routes sort_edges_1hop_* through the GPU partition kernel."""
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
    if edge_partition:
        for name in _DIST_NAMES:
            new = getattr(b2dist, name)
            for mod in ("anemoi.models.distributed.khop_edges", "anemoi.models.layers.block",
                        "anemoi.models.layers.mapper", "anemoi.models.layers.processor"):
                _rebind(mod, name, new)


def uninstall() -> None:
    for (module_name, attr), old in list(_saved.items()):
        setattr(importlib.import_module(module_name), attr, old)
    _saved.clear()
