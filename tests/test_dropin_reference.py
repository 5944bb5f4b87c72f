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


def _cpu_conv_patches(monkeypatch):
    """CPU stand-ins for the CUDA conv entry points (the oracle's restatements), so that the PLUGIN path -- the reference's own
    mapper / processor code driving this repo's blocks -- can be run end to end without a GPU."""
    import anemoi_models_b200.layers.conv as convmod
    from anemoi_models_b200 import ops
    from oracle import gtconv as og

    class CpuPlan:
        def __init__(self, edge_index, ns, nd):
            self.edge_index, self.num_src, self.num_dst, self.num_edges = edge_index, ns, nd, edge_index.shape[1]

    def cpu_conv(q, k, v, e, plan, halo=None):
        return og.gt_conv_unfused(q, k, v, e, plan.edge_index, (plan.num_src, plan.num_dst))

    def graphconv_forward(self, x, edge_attr, edge_index, size=None, plan=None):
        p = dict(self.named_parameters())
        return og.graph_conv_unfused(x, edge_attr, edge_index, {"edge_mlp." + k[len("edge_mlp."):]: v for k, v in p.items()},
                                     "edge_mlp.", size=size)

    monkeypatch.setattr(convmod, "get_csr", lambda ei, ns, nd: CpuPlan(ei, ns, nd))
    monkeypatch.setattr(ops, "gt_conv", cpu_conv)
    monkeypatch.setattr(convmod.GraphConv, "forward", graphconv_forward)


@pytest.mark.parametrize("kind", ["gt_forward_mapper", "gt_backward_mapper", "gt_processor", "gnn_processor", "gnn_forward_mapper",
                                  "gnn_backward_mapper"])
def test_reference_mappers_and_processors_give_the_same_numbers_through_the_plugin(reference_on_path, monkeypatch, kind):
    """The reference's mapper / processor classes, unchanged, once with the reference's blocks and once (after install()) with
    this repo's blocks, same weights, batch size 2: outputs, input gradients and every parameter gradient must agree.  (The conv
    arithmetic itself is the oracle's on the CPU here; on the GPU it is pinned by tests/test_gpu_*.py.)"""
    import anemoi_models_b200 as b2
    from anemoi.models.layers import mapper as ref_mapper
    from anemoi.models.layers import processor as ref_processor

    ns, nd, B, hid = 30, 20, 2, 32
    torch.manual_seed(3)
    bip, sq = _fake_graph(ns, nd, 70), _fake_graph(nd, nd, 60)
    rev = _fake_graph(nd, ns, 80)
    common = dict(trainable_size=6, sub_graph_edge_attributes=["edge_attr1"])
    if kind == "gt_forward_mapper":
        make = lambda: ref_mapper.GraphTransformerForwardMapper(in_channels_src=5, in_channels_dst=4, hidden_dim=hid, num_heads=4, sub_graph=bip,
                                                                src_grid_size=ns, dst_grid_size=nd, **common)
        x = (torch.randn(B * ns, 5), torch.randn(B * nd, 4))
    elif kind == "gt_backward_mapper":
        make = lambda: ref_mapper.GraphTransformerBackwardMapper(in_channels_src=hid, in_channels_dst=7, hidden_dim=hid, out_channels_dst=3, num_heads=4,
                                                                 sub_graph=rev, src_grid_size=nd, dst_grid_size=ns, **common)
        x = (torch.randn(B * nd, hid), torch.randn(B * ns, 7))
    elif kind == "gnn_forward_mapper":
        make = lambda: ref_mapper.GNNForwardMapper(in_channels_src=5, in_channels_dst=4, hidden_dim=hid, sub_graph=bip, src_grid_size=ns,
                                                   dst_grid_size=nd, **common)
        x = (torch.randn(B * ns, 5), torch.randn(B * nd, 4))
    elif kind == "gnn_backward_mapper":
        make = lambda: ref_mapper.GNNBackwardMapper(in_channels_src=hid, in_channels_dst=hid, hidden_dim=hid, out_channels_dst=3, sub_graph=rev,
                                                    src_grid_size=nd, dst_grid_size=ns, **common)
        x = (torch.randn(B * nd, hid), torch.randn(B * ns, hid))  # the GNN backward mapper does not embed its dst input
    elif kind == "gt_processor":
        make = lambda: ref_processor.GraphTransformerProcessor(num_layers=2, num_channels=hid, num_chunks=1, num_heads=4, sub_graph=sq,
                                                               src_grid_size=nd, dst_grid_size=nd, **common)
        x = torch.randn(B * nd, hid)
    else:
        make = lambda: ref_processor.GNNProcessor(num_layers=2, num_channels=hid, num_chunks=1, sub_graph=sq, src_grid_size=nd,
                                                  dst_grid_size=nd, **common)
        x = torch.randn(B * nd, hid)
    pair = isinstance(x, tuple)
    shapes = ([list(x[0].shape)], [list(x[1].shape)]) if pair else ([list(x.shape)],)
    if not pair:
        shapes = (shapes[0], shapes[0])

    def run(mod):
        xin = tuple(t.clone().requires_grad_(True) for t in x) if pair else x.clone().requires_grad_(True)
        out = mod(xin, batch_size=B, shard_shapes=shapes)
        outs = [o for o in (out if isinstance(out, tuple) else (out,)) if o.requires_grad]
        gen = torch.Generator().manual_seed(4)
        sum((o * torch.randn(o.shape, generator=gen)).sum() for o in outs).backward()
        grads_in = [t.grad for t in (xin if pair else (xin,))]
        return [o.detach() for o in outs], grads_in, {n: p.grad for n, p in mod.named_parameters()}

    ref_mod = make()
    ref_out, ref_gin, ref_gp = run(ref_mod)
    b2.install(edge_partition=False)  # the partition helper is a CUDA kernel; on one rank it only regroups the edges
    _cpu_conv_patches(monkeypatch)
    new_mod = make()
    assert any(type(m).__module__.startswith("anemoi_models_b200") for m in new_mod.modules())
    new_mod.load_state_dict(ref_mod.state_dict())
    new_out, new_gin, new_gp = run(new_mod)
    assert len(new_out) == len(ref_out)
    for a, b in zip(new_out, ref_out):
        assert torch.allclose(a, b, atol=2e-5), float((a - b).abs().max())
    for a, b in zip(new_gin, ref_gin):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.allclose(a, b, atol=2e-5 * max(1.0, float(b.abs().max())))
    for n, g in ref_gp.items():
        if g is None:
            assert new_gp[n] is None or float(new_gp[n].abs().max()) == 0.0, n
        else:
            assert new_gp[n] is not None and torch.allclose(new_gp[n], g, atol=5e-5 * max(1.0, float(g.abs().max()))), n


class _Idx:
    def __init__(self, n, prognostic, full=None, diagnostic=(), names=None):
        self._n, self.prognostic, self.full, self.diagnostic = n, list(prognostic), list(full if full is not None else range(n)), list(diagnostic)
        self.name_to_index = names or {f"v{i}": i for i in range(n)}

    def __len__(self):
        return self._n


def _tiny_model_inputs(kind):
    from types import SimpleNamespace

    from anemoi.utils.config import DotDict
    from torch_geometric.data import HeteroData

    gen = torch.Generator().manual_seed(21)
    n_data, n_hid = 40, 16
    g = HeteroData()
    g["data"].x = torch.rand(n_data, 2, generator=gen) * 3 - 1.5
    g["hidden"].x = torch.rand(n_hid, 2, generator=gen) * 3 - 1.5
    for (a, na), (b, nb), e in ((("data", n_data), ("hidden", n_hid), 90), (("hidden", n_hid), ("hidden", n_hid), 64), (("hidden", n_hid), ("data", n_data), 120)):
        st = g[(a, "to", b)]
        st.edge_index = torch.stack([torch.randint(0, na, (e,), generator=gen), torch.randint(0, nb, (e,), generator=gen)])
        st.edge_length = torch.rand(e, 1, generator=gen)
        st.edge_dirs = torch.rand(e, 2, generator=gen)
    attrs = ["edge_length", "edge_dirs"]
    p = "anemoi.models.layers."
    if kind == "graphtransformer":
        enc = {"_target_": p + "mapper.GraphTransformerForwardMapper", "trainable_size": 3, "sub_graph_edge_attributes": attrs, "num_chunks": 1,
               "num_heads": 4, "mlp_hidden_ratio": 2, "activation": "GELU"}
        proc = {"_target_": p + "processor.GraphTransformerProcessor", "trainable_size": 3, "sub_graph_edge_attributes": attrs, "num_layers": 2,
                "num_chunks": 2, "num_heads": 4, "mlp_hidden_ratio": 2, "activation": "GELU"}
        dec = dict(enc, _target_=p + "mapper.GraphTransformerBackwardMapper")
    else:
        enc = {"_target_": p + "mapper.GNNForwardMapper", "trainable_size": 3, "sub_graph_edge_attributes": attrs, "num_chunks": 1,
               "mlp_extra_layers": 0, "activation": "SiLU"}
        proc = {"_target_": p + "processor.GNNProcessor", "trainable_size": 3, "sub_graph_edge_attributes": attrs, "num_layers": 2,
                "num_chunks": 1, "mlp_extra_layers": 0, "activation": "SiLU"}
        dec = dict(enc, _target_=p + "mapper.GNNBackwardMapper")
    cfg = DotDict({"graph": {"data": "data", "hidden": "hidden"}, "training": {"multistep_input": 2},
                   "model": {"num_channels": 32, "trainable_parameters": {"data": 2, "hidden": 2}, "encoder": enc, "processor": proc,
                             "decoder": dec, "bounding": []}})
    nvar = 4
    idx = SimpleNamespace(internal_model=SimpleNamespace(input=_Idx(nvar, [0, 1, 2]), output=_Idx(nvar, [0, 1, 2], diagnostic=[3])))
    x = torch.randn(2, 2, 1, n_data, nvar, generator=gen)  # batch, time, ensemble, grid, vars
    return cfg, idx, g, x


@pytest.mark.parametrize("kind", ["graphtransformer", "gnn"])
def test_full_reference_model_runs_unchanged_on_the_plugin(reference_on_path, monkeypatch, kind):
    """`AnemoiModelEncProcDec` (reference models/encoder_processor_decoder.py, unmodified; hydra / anemoi-utils stood in for by the
    10-line shims of oracle/ref_shims): built once from the reference's blocks and once after `install()`, same state_dict --
    forward output (batch 2, activation checkpointing as the model does it) and every parameter gradient agree."""
    shims = os.path.join(ROOT, "oracle", "ref_shims")
    monkeypatch.syspath_prepend(shims)
    import anemoi_models_b200 as b2
    from anemoi.models.models.encoder_processor_decoder import AnemoiModelEncProcDec

    cfg, idx, graph, x = _tiny_model_inputs(kind)
    torch.manual_seed(0)
    ref = AnemoiModelEncProcDec(model_config=cfg, data_indices=idx, graph_data=graph)
    ref_out = ref(x)
    gen = torch.Generator().manual_seed(1)
    w = torch.randn(ref_out.shape, generator=gen)
    (ref_out * w).sum().backward()

    b2.install(edge_partition=False)
    _cpu_conv_patches(monkeypatch)
    new = AnemoiModelEncProcDec(model_config=cfg, data_indices=idx, graph_data=graph)
    assert sum(type(m).__module__.startswith("anemoi_models_b200.layers.block") for m in new.modules()) >= 4
    assert _state(new) == _state(ref)
    new.load_state_dict(ref.state_dict())
    new_out = new(x)
    (new_out * w).sum().backward()
    assert new_out.shape == ref_out.shape and torch.allclose(new_out, ref_out, atol=3e-5), float((new_out - ref_out).abs().max())
    ref_g = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in new.named_parameters():
        g = ref_g[n]
        if g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
        else:
            assert p.grad is not None and torch.allclose(p.grad, g, atol=1e-4 * max(1.0, float(g.abs().max()))), n


def test_hierarchical_reference_model_runs_unchanged_on_the_plugin(reference_on_path, monkeypatch):
    """`AnemoiModelEncProcDecHierarchical` (reference models/hierarchical.py:30-308, unmodified): two hidden levels with level
    processors, down- and up-scale mappers -- the same kernels at more call sites."""
    from types import SimpleNamespace

    monkeypatch.syspath_prepend(os.path.join(ROOT, "oracle", "ref_shims"))
    import anemoi_models_b200 as b2
    from anemoi.models.models.hierarchical import AnemoiModelEncProcDecHierarchical
    from anemoi.utils.config import DotDict
    from torch_geometric.data import HeteroData

    gen = torch.Generator().manual_seed(33)
    sizes = {"data": 36, "hidden_1": 18, "hidden_2": 8}
    g = HeteroData()
    for name, n in sizes.items():
        g[name].x = torch.rand(n, 2, generator=gen) * 3 - 1.5
    for a, b, e in (("data", "hidden_1", 80), ("hidden_1", "hidden_1", 60), ("hidden_2", "hidden_2", 30), ("hidden_1", "hidden_2", 40),
                    ("hidden_2", "hidden_1", 50), ("hidden_1", "data", 100)):
        st = g[(a, "to", b)]
        st.edge_index = torch.stack([torch.randint(0, sizes[a], (e,), generator=gen), torch.randint(0, sizes[b], (e,), generator=gen)])
        st.edge_length = torch.rand(e, 1, generator=gen)
    p = "anemoi.models.layers."
    enc = {"_target_": p + "mapper.GraphTransformerForwardMapper", "trainable_size": 2, "sub_graph_edge_attributes": ["edge_length"],
           "num_chunks": 1, "num_heads": 4, "mlp_hidden_ratio": 2, "activation": "GELU"}
    proc = {"_target_": p + "processor.GraphTransformerProcessor", "trainable_size": 2, "sub_graph_edge_attributes": ["edge_length"],
            "num_layers": 2, "num_chunks": 1, "num_heads": 4, "mlp_hidden_ratio": 2, "activation": "GELU"}
    cfg = DotDict({"graph": {"data": "data", "hidden": ["hidden_1", "hidden_2"]}, "training": {"multistep_input": 1},
                   "model": {"num_channels": 16, "trainable_parameters": {"hidden": 2}, "enable_hierarchical_level_processing": True,
                             "level_process_num_layers": 1, "encoder": enc, "processor": proc,
                             "decoder": dict(enc, _target_=p + "mapper.GraphTransformerBackwardMapper"), "bounding": []}})
    idx = SimpleNamespace(internal_model=SimpleNamespace(input=_Idx(3, [0, 1]), output=_Idx(3, [0, 1], diagnostic=[2])))
    x = torch.randn(2, 1, 1, sizes["data"], 3, generator=gen)

    torch.manual_seed(0)
    ref = AnemoiModelEncProcDecHierarchical(model_config=cfg, data_indices=idx, graph_data=g)
    ref_out = ref(x)
    w = torch.randn(ref_out.shape, generator=torch.Generator().manual_seed(2))
    (ref_out * w).sum().backward()
    b2.install(edge_partition=False)
    _cpu_conv_patches(monkeypatch)
    new = AnemoiModelEncProcDecHierarchical(model_config=cfg, data_indices=idx, graph_data=g)
    assert sum(type(m).__module__.startswith("anemoi_models_b200.layers.block") for m in new.modules()) >= 7
    assert _state(new) == _state(ref)
    new.load_state_dict(ref.state_dict())
    new_out = new(x)
    (new_out * w).sum().backward()
    assert torch.allclose(new_out, ref_out, atol=3e-5), float((new_out - ref_out).abs().max())
    ref_g = {n: p_.grad for n, p_ in ref.named_parameters()}
    for n, p_ in new.named_parameters():
        gref = ref_g[n]
        if gref is None:
            assert p_.grad is None or float(p_.grad.abs().max()) == 0.0, n
        else:
            assert p_.grad is not None and torch.allclose(p_.grad, gref, atol=1e-4 * max(1.0, float(gref.abs().max()))), n


@pytest.mark.parametrize("kind", ["gt_forward_mapper", "gt_processor", "gnn_processor"])
def test_plugin_forward_under_inference_mode(reference_on_path, monkeypatch, kind):
    """Lightning's validate / predict loops and inference runners call the model under `torch.inference_mode()`: tensors made
    there do not track a version counter (reading `_version` raises).  The memoised `_expand_edges` and the plan cache must
    work there, give the numbers of a `no_grad` forward, and leave tensors a later TRAINING forward can use."""
    import anemoi_models_b200 as b2
    from anemoi.models.layers import mapper as ref_mapper
    from anemoi.models.layers import processor as ref_processor
    from anemoi_models_b200.graph import TensorKeyedCache

    ns, nd, B, hid = 30, 20, 2, 32
    torch.manual_seed(5)
    common = dict(trainable_size=6, sub_graph_edge_attributes=["edge_attr1"])
    b2.install(edge_partition=False)
    _cpu_conv_patches(monkeypatch)
    if kind == "gt_forward_mapper":
        mod = ref_mapper.GraphTransformerForwardMapper(in_channels_src=5, in_channels_dst=4, hidden_dim=hid, num_heads=4,
                                                       sub_graph=_fake_graph(ns, nd, 70), src_grid_size=ns, dst_grid_size=nd, **common)
        x = (torch.randn(B * ns, 5), torch.randn(B * nd, 4))
        shapes = ([list(x[0].shape)], [list(x[1].shape)])
    else:
        cls = ref_processor.GraphTransformerProcessor if kind == "gt_processor" else ref_processor.GNNProcessor
        extra = dict(num_heads=4) if kind == "gt_processor" else {}
        mod = cls(num_layers=2, num_channels=hid, num_chunks=1, sub_graph=_fake_graph(nd, nd, 60), src_grid_size=nd, dst_grid_size=nd,
                  **extra, **common)
        x = torch.randn(B * nd, hid)
        shapes = ([list(x.shape)], [list(x.shape)])

    def first(o):
        return o[1] if isinstance(o, tuple) else o

    with torch.inference_mode():
        a1 = first(mod(x, batch_size=B, shard_shapes=shapes))
        a2 = first(mod(x, batch_size=B, shard_shapes=shapes))  # second call: cache hits
    with torch.no_grad():
        ref = first(mod(x, batch_size=B, shard_shapes=shapes))
    assert torch.equal(a1, a2) and torch.allclose(a1, ref, atol=1e-6)
    # a training forward + backward afterwards still works (nothing cached is an inference tensor that autograd would reject)
    out = first(mod(x, batch_size=B, shard_shapes=shapes))
    out.sum().backward()
    assert any(p.grad is not None for p in mod.parameters())
    # the plan cache itself, keyed by an inference tensor
    cache = TensorKeyedCache()
    with torch.inference_mode():
        ei = torch.cat([torch.randint(0, 5, (2, 7)), torch.randint(0, 5, (2, 3))], dim=1)
        assert ei.is_inference()
        v1 = cache.get(ei, ("k",), lambda: object())
        assert cache.get(ei, ("k",), lambda: object()) is v1
        assert cache.get(ei.clone(), ("k",), lambda: object()) is v1  # same content, other object
