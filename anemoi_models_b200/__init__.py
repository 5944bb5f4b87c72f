"""anemoi_models_b200 -- B200-native (sm_100a) graph message-passing hot path for ecmwf/anemoi-models.

Drop-in replacements for `anemoi.models.layers.conv.{GraphTransformerConv, GraphConv}` and the blocks that
use them (`anemoi.models.layers.block.Graph*Block`), backed by hand-written CUDA kernels behind a C ABI
(`include/anemoi_b200.h`, `lib/libanemoi_b200.so`).  CUDA only: there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"

from .graph import GraphCSR, clear_csr_cache, get_csr  # noqa: F401
from .layers.conv import GraphConv, GraphTransformerConv  # noqa: F401
from .layers.block import (  # noqa: F401
    GraphConvMapperBlock,
    GraphConvProcessorBlock,
    GraphTransformerMapperBlock,
    GraphTransformerProcessorBlock,
)
from .install import install, uninstall  # noqa: F401
