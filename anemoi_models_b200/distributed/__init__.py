"""dst-node sharding with a halo exchange: replacement for `anemoi.models.distributed` on the graph path."""
from .shapes import change_channels_in_shape, get_shape_shards  # noqa: F401
from .khop_edges import get_k_hop_edges, sort_edges_1hop_chunks, sort_edges_1hop_sharding  # noqa: F401
from .collectives import gather_tensor, reduce_shard_tensor, reduce_tensor, shard_tensor, sync_tensor  # noqa: F401
from .halo import (HaloPlan, build_bipartite_halo_plan, build_local_halo_plan, exchange_rows, halo_exchange,  # noqa: F401
                   halo_gather, return_rows, select_sharded_edges)
from .halo import aligned_bounds_from_ranges, aligned_src_bounds  # noqa: F401
from .transformer import shard_heads, shard_sequence  # noqa: F401
