"""Drop-in check against the REAL reference package (only in the build container, where /root/reference exists; the GPU
box skips it).  After `install()` the reference's own mappers / processors, constructed through the reference's code,
must be built from this repo's blocks and convs, with exactly the reference's parameter names and shapes -- so a
reference checkpoint loads unchanged."""
import os
import sys

import pytest
import torch

from conftest import ROOT

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")


@pytest.fixture()
def reference_on_path():
    added = [p for p in (os.path.join(ROOT, "oracle", "pyg_shim"), REF_SRC) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    yield
    import anemoi_models_b200 as b2

    b2.uninstall()
    for p in added:
        sys.path.remove(p)


def _fake_graph(ns, nd, e):
    from torch_geometric.data import HeteroData

    g = HeteroData()
    g[("src", "to", "dst")].edge_index = torch.stack([torch.randint(0, ns, (e,)), torch.randint(0, nd, (e,))])
    g[("src", "to", "dst")].edge_attr1 = torch.rand(e, 3)
    return g[("src", "to", "dst")]


def _state(mod):
    return {k: tuple(v.shape) for k, v in mod.state_dict().items()}


def test_reference_modules_pick_up_the_b200_path(reference_on_path):
    import anemoi_models_b200 as b2
    from anemoi.models.layers import mapper as ref_mapper
    from anemoi.models.layers import processor as ref_processor

    torch.manual_seed(0)
    sub = _fake_graph(30, 20, 60)
    kw = dict(in_channels_src=5, in_channels_dst=4, hidden_dim=32, trainable_size=6, num_heads=4, sub_graph=sub,
              sub_graph_edge_attributes=["edge_attr1"], src_grid_size=30, dst_grid_size=20)
    ref_enc = ref_mapper.GraphTransformerForwardMapper(**kw)
    ref_proc = ref_processor.GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, trainable_size=6,
                                                       sub_graph=_fake_graph(20, 20, 50), sub_graph_edge_attributes=["edge_attr1"],
                                                       src_grid_size=20, dst_grid_size=20)
    ref_gnn = ref_processor.GNNProcessor(num_layers=2, num_channels=32, num_chunks=1, trainable_size=6,
                                         sub_graph=_fake_graph(20, 20, 50), sub_graph_edge_attributes=["edge_attr1"],
                                         src_grid_size=20, dst_grid_size=20)
    b2.install()
    new_enc = ref_mapper.GraphTransformerForwardMapper(**kw)
    new_proc = ref_processor.GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, trainable_size=6,
                                                       sub_graph=_fake_graph(20, 20, 50), sub_graph_edge_attributes=["edge_attr1"],
                                                       src_grid_size=20, dst_grid_size=20)
    new_gnn = ref_processor.GNNProcessor(num_layers=2, num_channels=32, num_chunks=1, trainable_size=6,
                                         sub_graph=_fake_graph(20, 20, 50), sub_graph_edge_attributes=["edge_attr1"],
                                         src_grid_size=20, dst_grid_size=20)
    # built from this repo's classes ...
    assert isinstance(new_enc.proc, b2.GraphTransformerMapperBlock) and isinstance(new_enc.proc.conv, b2.GraphTransformerConv)
    blocks = [m for m in new_proc.modules() if isinstance(m, b2.GraphTransformerProcessorBlock)]
    assert len(blocks) == 2 and all(isinstance(b.conv, b2.GraphTransformerConv) for b in blocks)
    gblocks = [m for m in new_gnn.modules() if isinstance(m, b2.GraphConvProcessorBlock)]
    assert len(gblocks) == 2 and all(isinstance(b.conv, b2.GraphConv) for b in gblocks)
    # ... with the reference's parameter names and shapes: checkpoints are interchangeable
    assert _state(new_enc) == _state(ref_enc)
    assert _state(new_proc) == _state(ref_proc)
    assert _state(new_gnn) == _state(ref_gnn)
    new_enc.load_state_dict(ref_enc.state_dict())
    new_gnn.load_state_dict(ref_gnn.state_dict())
    # constant edge buffers: the expanded edge_index is the same tensor object on every forward (plan cache hits by identity)
    e1 = new_enc._expand_edges(new_enc.edge_index_base, new_enc.edge_inc, 2)
    e2 = new_enc._expand_edges(new_enc.edge_index_base, new_enc.edge_inc, 2)
    assert e1 is e2 and torch.equal(e1, ref_enc._expand_edges(ref_enc.edge_index_base, ref_enc.edge_inc, 2))
    assert new_enc._expand_edges(new_enc.edge_index_base, new_enc.edge_inc, 1) is not e1
    # the GPU partition helper is bound where the reference looks it up
    import anemoi.models.layers.block as ref_block
    from anemoi_models_b200.distributed import sort_edges_1hop_chunks

    assert ref_block.sort_edges_1hop_chunks is sort_edges_1hop_chunks
    b2.uninstall()
    from anemoi.models.layers.conv import GraphTransformerConv as RefConv

    assert ref_block.GraphTransformerConv is RefConv and not isinstance(ref_mapper.GraphTransformerForwardMapper(**kw).proc, b2.GraphTransformerMapperBlock)


def test_index_helpers_equal_the_reference_functions(reference_on_path):
    """get_k_hop_edges, get_shape_shards / change_channels_in_shape, and the single-rank reshapes of shard_qkve_heads /
    shard_output_seq against the reference's own functions (imported from /root/reference, PyG through the oracle's shim)."""
    import anemoi_models_b200 as b2
    from anemoi.models.distributed import khop_edges as ref_khop
    from anemoi.models.distributed import shapes as ref_shapes
    from anemoi.models.layers import block as ref_block
    from anemoi_models_b200 import distributed as b2dist

    gen = torch.Generator().manual_seed(11)
    ei = torch.randint(0, 25, (2, 90), generator=gen)
    ea = torch.randn(90, 4, generator=gen)
    nodes = torch.tensor([0, 7, 8, 21])
    for hops in (1, 2):
        ref_ea, ref_ei = ref_khop.get_k_hop_edges(nodes, ea, ei, num_hops=hops)
        got_ea, got_ei = b2dist.get_k_hop_edges(nodes, ea, ei, num_hops=hops)
        assert torch.equal(ref_ea, got_ea) and torch.equal(ref_ei, got_ei)
    x = torch.empty(23, 6)
    assert ref_shapes.get_shape_shards(x, 0, None) == b2dist.get_shape_shards(x, 0, None)
    sl = [[5, 6], [4, 6]]
    assert ref_shapes.change_channels_in_shape(sl, 9) == b2dist.change_channels_in_shape(sl, 9)

    torch.manual_seed(0)
    kw = dict(in_channels=24, hidden_dim=16, out_channels=24, edge_dim=3, num_heads=4)
    ref_blk, new_blk = ref_block.GraphTransformerProcessorBlock(**kw), b2.GraphTransformerProcessorBlock(**kw)
    q, k, v, e = (torch.randn(n, 24, generator=gen) for n in (10, 14, 14, 30))  # batch_size 2: 5 / 7 / 15 rows per sample
    shapes = ([[14, 24]], [[10, 24]], [[30, 24]])
    ref_out = ref_blk.shard_qkve_heads(q, k, v, e, shapes, 2, None)
    new_out = new_blk.shard_qkve_heads(q, k, v, e, shapes, 2, None)
    for a, b in zip(ref_out, new_out):
        assert torch.equal(a, b)
    assert torch.equal(ref_blk.shard_output_seq(ref_out[0], shapes, 2, None), new_blk.shard_output_seq(new_out[0], shapes, 2, None))
