"""Host-side mirror of the reference interface: constructors, attribute / state_dict names, argument checking,
shape bookkeeping, plan caching -- everything that does not need a GPU.  CPU tensors must be refused loudly."""
import numpy as np
import pytest
import torch

import anemoi_models_b200 as b2
from anemoi_models_b200.distributed import shapes as b2shapes
from anemoi_models_b200.graph import TensorKeyedCache, check_edge_index, resolve_size
from conftest import load_golden, t


def test_conv_constructor_and_attributes():
    conv = b2.GraphTransformerConv(out_channels=16)
    assert conv.out_channels == 16 and conv.dropout == 0.0
    assert list(conv.state_dict()) == []
    gc = b2.GraphConv(in_channels=8, out_channels=8, mlp_extra_layers=1)
    keys = list(gc.state_dict())
    assert keys[0] == "edge_mlp.model.0.weight" and gc.edge_mlp.model[0].in_features == 24
    assert "edge_mlp.model.7.weight" in keys  # Linear,act,(Linear,act)x2,Linear,LayerNorm -> LN is index 7
    with pytest.raises(TypeError):
        b2.GraphTransformerConv(out_channels=4, bogus=1)
    with pytest.raises(RuntimeError):
        b2.GraphConv(4, 4, activation="NotAnActivation")


@pytest.mark.parametrize("name,cls", [("block_gt_mapper.npz", "GraphTransformerMapperBlock"),
                                      ("block_gt_processor.npz", "GraphTransformerProcessorBlock")])
def test_gt_block_state_dict_names_match_reference(name, cls):
    z = load_golden(name)
    _, _, D, H, ed, hid = (int(x) for x in z["meta"])
    blk = getattr(b2, cls)(in_channels=D, hidden_dim=hid, out_channels=D, edge_dim=ed, num_heads=H)
    ref_keys = sorted(k[2:] for k in z if k.startswith("p."))
    assert sorted(blk.state_dict()) == ref_keys
    blk.load_state_dict({k[2:]: t(v) for k, v in z.items() if k.startswith("p.")})
    # attributes the reference's tests assert on (tests/layers/block/test_block_graphtransformer.py:84-101)
    assert blk.out_channels_conv == D // H and blk.num_heads == H and blk.num_chunks == 1
    assert isinstance(blk.conv, b2.GraphTransformerConv) and isinstance(blk.lin_key, torch.nn.Linear)
    q = torch.zeros(7, D)
    qs, ks, vs, es = blk.shard_qkve_heads(q, q, q, q, None, 1)
    assert qs.shape == (7, H, D // H)
    assert blk.shard_output_seq(qs, None, 1).shape == (7, D)


@pytest.mark.parametrize("name,cls", [("block_graphconv_processor.npz", "GraphConvProcessorBlock"),
                                      ("block_graphconv_mapper.npz", "GraphConvMapperBlock")])
def test_graphconv_block_state_dict_names_match_reference(name, cls):
    z = load_golden(name)
    D = int(z["meta"][-1])
    blk = getattr(b2, cls)(in_channels=D, out_channels=D)
    assert sorted(blk.state_dict()) == sorted(k[2:] for k in z if k.startswith("p."))
    assert blk.update_src_nodes is True and blk.num_chunks == 1


def test_unknown_activation_raises_runtime_error():
    with pytest.raises(RuntimeError):
        b2.GraphTransformerProcessorBlock(8, 16, 8, edge_dim=3, num_heads=2, activation="Nope")


def test_edge_index_and_size_checks_like_pyg():
    with pytest.raises(ValueError):
        check_edge_index(torch.zeros(2, 3))  # float
    with pytest.raises(ValueError):
        check_edge_index(torch.zeros(3, 3, dtype=torch.long))
    with pytest.raises(ValueError):
        check_edge_index(torch.zeros(6, dtype=torch.long))
    with pytest.raises(ValueError):
        check_edge_index([[0], [1]])
    assert resolve_size(None, 5, 3) == (5, 3)
    assert resolve_size((None, 3), 5, 3) == (5, 3)
    with pytest.raises(ValueError):
        resolve_size((4, 3), 5, 3)
    with pytest.raises(ValueError):
        resolve_size((5, 2), 5, 3)


def test_cpu_tensors_are_refused_loudly():
    z = load_golden("gtconv_kat3.npz")
    conv = b2.GraphTransformerConv(out_channels=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), (2, 3))
    with pytest.raises(TypeError):
        conv(t(z["q"]), t(z["k"]), t(z["v"]), None, t(z["edge_index"]), (2, 3))
    with pytest.raises(ValueError):
        conv(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), (7, 3))
    gc = b2.GraphConv(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        gc(torch.zeros(3, 4), torch.zeros(2, 4), torch.tensor([[0, 1], [1, 2]]))


def test_shape_shards_bit_exact():
    z = load_golden("sharding.npz")
    for key, (rows, P) in {"shards_10_3": (10, 3), "shards_40320_8": (40320, 8), "shards_7_8": (7, 8)}.items():
        sizes = b2shapes.tensor_split_sizes(rows, P)
        assert sizes == [int(s[0]) for s in z[key]]
        assert sizes == [x.shape[0] for x in torch.tensor_split(torch.empty(rows), P)]
    assert b2shapes.get_shape_shards(torch.empty(9, 4), 0, None) == [[9, 4]]
    assert b2shapes.change_channels_in_shape([[5, 3], [4, 3]], 7) == [[5, 7], [4, 7]]
    assert b2shapes.change_channels_in_shape([], 7) == []
    assert b2shapes.bounds_from_shapes([[5, 3], [4, 3], [4, 3]]) == [0, 5, 9, 13]


def test_tensor_keyed_cache_identity_content_and_version():
    cache = TensorKeyedCache(maxsize=2)
    built = []

    def builder(tag):
        def f():
            built.append(tag)
            return tag
        return f

    a = torch.tensor([[0, 1], [1, 0]])
    assert cache.get(a, (2, 2), builder("A")) == "A"
    assert cache.get(a, (2, 2), builder("A2")) == "A" and built == ["A"]          # identity hit
    assert cache.get(a.clone(), (2, 2), builder("A3")) == "A" and built == ["A"]  # content hit
    assert cache.get(a, (3, 2), builder("B")) == "B"                             # other sizes -> new plan
    a[0, 0] = 1                                                                   # in-place edit bumps _version
    assert cache.get(a, (2, 2), builder("C")) == "C"
    assert built == ["A", "B", "C"]


def test_install_rebinds_and_restores_reference_symbols():
    import sys
    import types

    # a stand-in for the reference package layout (the real one is only present in the build container)
    made = []
    for name in ("anemoi", "anemoi.models", "anemoi.models.layers", "anemoi.models.layers.block",
                 "anemoi.models.layers.conv", "anemoi.models.layers.mapper", "anemoi.models.layers.chunk"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            made.append(name)
    blockmod = sys.modules["anemoi.models.layers.block"]
    sentinel = object()
    old = getattr(blockmod, "GraphTransformerConv", None)
    blockmod.GraphTransformerConv = sentinel
    try:
        b2.install()
        assert blockmod.GraphTransformerConv is b2.GraphTransformerConv
        b2.uninstall()
        assert blockmod.GraphTransformerConv is sentinel
    finally:
        if old is None:
            del blockmod.GraphTransformerConv
        else:
            blockmod.GraphTransformerConv = old
        for name in made:
            sys.modules.pop(name, None)


@pytest.mark.parametrize("hops", [1, 2, 3])
def test_get_k_hop_edges_matches_pyg_semantics(hops):
    """reference khop_edges.py:24-47 over PyG's k_hop_subgraph(directed=True) (restated in oracle/pyg_shim): same edges, same order."""
    from torch_geometric.utils import k_hop_subgraph, mask_to_index  # the oracle's shim (tests/conftest.py puts it on the path)

    from anemoi_models_b200.distributed import get_k_hop_edges

    gen = torch.Generator().manual_seed(hops)
    n, E = 40, 150
    ei = torch.randint(0, n, (2, E), generator=gen)
    ea = torch.randn(E, 3, generator=gen)
    nodes = torch.tensor([3, 4, 5, 17])
    _, ei_ref, _, mask = k_hop_subgraph(node_idx=nodes, num_hops=hops, edge_index=ei, directed=True)
    got_ea, got_ei = get_k_hop_edges(nodes, ea, ei, num_hops=hops)
    assert torch.equal(got_ei, ei_ref) and torch.equal(got_ea, ea[mask_to_index(mask)])
    with pytest.raises(ValueError):
        get_k_hop_edges(nodes, ea, ei, num_hops=0)


def test_folded_lin_edge_host_glue_against_reference(monkeypatch):
    """Round-2 draft (ops.gt_conv_folded): the host side -- padding with the bias column, the [Nd]-sized einsums with W, the
    assembly of dq / dW / db / d raw -- checked on the CPU with a torch emulation standing in for the three CUDA kernel calls,
    against the reference op sequence conv(q, k, v, lin_edge(raw)) with autograd."""
    import math
    from types import SimpleNamespace

    from anemoi_models_b200 import ops
    from oracle import gtconv as og

    gen = torch.Generator().manual_seed(9)
    ns, nd, E, H, C, ed = 31, 17, 140, 4, 8, 11
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd - 1, (E,), generator=gen)])
    src, dst = ei[0], ei[1]
    state = {}

    def seg_sum(x, index, n):
        return x.new_zeros((n,) + tuple(x.shape[1:])).index_add_(0, index, x)

    def emul_fwd(q, k, v, rawp, qw, plan):
        s = ((q[dst] * k[src]).sum(-1) + (qw[dst] * rawp[:, None, :]).sum(-1)) / math.sqrt(C)
        a = og.segment_softmax(s, dst, nd)
        state["a"] = a
        return seg_sum(a[..., None] * v[src], dst, nd), torch.zeros(nd, H), seg_sum(a[..., None] * rawp[:, None, :], dst, nd)

    def emul_bwd(q, k, v, rawp, qw, gw, out, lse2, g, plan):
        a = state["a"]
        Dl = (g * out).sum(-1)
        gv = (g[dst] * v[src]).sum(-1) + (gw[dst] * rawp[:, None, :]).sum(-1)
        dss = a * (gv - Dl[dst]) / math.sqrt(C)
        draw = (a[..., None] * gw[dst]).sum(1) + (dss[..., None] * qw[dst]).sum(1)
        return (seg_sum(dss[..., None] * k[src], dst, nd), seg_sum(dss[..., None] * rawp[:, None, :], dst, nd),
                seg_sum(dss[..., None] * q[dst], src, ns), seg_sum(a[..., None] * g[dst], src, ns), draw)

    monkeypatch.setattr(ops, "_fold_fwd_kernel", emul_fwd)
    monkeypatch.setattr(ops, "_fold_bwd_kernels", emul_bwd)
    for with_bias in (True, False):
        q, k, v = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns))
        raw = torch.rand(E, ed, generator=gen)
        lin = torch.nn.Linear(ed, H * C, bias=with_bias)
        g = torch.randn(nd, H, C, generator=gen)
        ref_in = [t.clone().requires_grad_(True) for t in (q, k, v, raw)]
        ref = og.gt_conv_unfused(ref_in[0], ref_in[1], ref_in[2], lin(ref_in[3]).view(E, H, C), ei, (ns, nd))
        ref.backward(g)
        ref_grads = [t.grad for t in ref_in] + [lin.weight.grad.clone()] + ([lin.bias.grad.clone()] if with_bias else [])
        lin.zero_grad()
        got_in = [t.clone().requires_grad_(True) for t in (q, k, v, raw)]
        out = ops._GTConvFoldedFn.apply(*got_in, lin.weight, lin.bias, SimpleNamespace(num_edges=E, num_src=ns, num_dst=nd))
        out.backward(g)
        got_grads = [t.grad for t in got_in] + [lin.weight.grad] + ([lin.bias.grad] if with_bias else [])
        assert torch.allclose(out, ref, atol=2e-5)
        for a_, b_ in zip(got_grads, ref_grads):
            assert a_.shape == b_.shape and float((a_ - b_).abs().max()) <= 2e-5 * max(1.0, float(b_.abs().max()))
    with pytest.raises(ValueError):
        ops._fold_pad(torch.zeros(3, 16), torch.zeros(8, 16), None, 2, 4)


@pytest.mark.parametrize("fixture,kind", [("block_gt_mapper.npz", "mapper"), ("block_gt_processor.npz", "processor")])
def test_block_folded_branch_plumbing_against_golden(monkeypatch, fixture, kind):
    """The blocks' AB2_EDGE_FOLD branch (round-2 switch) on the CPU: `ops.gt_conv_folded` is replaced by the reference op sequence
    on lin_edge(raw), so what is checked is the branch itself -- argument order, reshapes, size resolution, that lin_edge's
    parameters and the raw edge features still get their gradients -- against the golden vectors of the reference block."""
    import torch.nn.functional as F

    from anemoi_models_b200 import ops
    from anemoi_models_b200.layers import block as b2block
    from oracle import gtconv as og

    class FakePlan:
        def __init__(self, edge_index, ns, nd):
            self.edge_index, self.num_src, self.num_dst, self.num_edges = edge_index, ns, nd, edge_index.shape[1]

    def folded_cpu(q, k, v, raw, weight, bias, plan):
        E, (H, C) = raw.shape[0], q.shape[1:]
        return og.gt_conv_unfused(q, k, v, F.linear(raw, weight, bias).view(E, H, C), plan.edge_index, (plan.num_src, plan.num_dst))

    monkeypatch.setattr(b2block, "_fold_applies", lambda query, edge_attr: True)
    monkeypatch.setattr(b2block, "get_csr", lambda ei, ns, nd: FakePlan(ei, ns, nd))
    monkeypatch.setattr(ops, "gt_conv_folded", folded_cpu)
    z = load_golden(fixture)
    ns, nd, D, H, ed, hid = (int(x) for x in z["meta"])
    cls = b2.GraphTransformerMapperBlock if kind == "mapper" else b2.GraphTransformerProcessorBlock
    blk = cls(D, hid, D, edge_dim=ed, num_heads=H)
    blk.load_state_dict({k[2:]: t(v) for k, v in z.items() if k.startswith("p.")})
    ei, ea = t(z["edge_index"]), t(z["ea"]).requires_grad_(True)
    if kind == "mapper":
        xs, xd = t(z["xs"]).requires_grad_(True), t(z["xd"]).requires_grad_(True)
        (_, out), _ = blk((xs, xd), ea, ei, ([[ns, D]], [[nd, D]], [[ea.shape[0], ed]]), 1, size=(ns, nd))
        ref_out, grads = t(z["dst_new"]), [(xs, "dxs"), (xd, "dxd"), (ea, "dea")]
    else:
        xd = t(z["x"]).requires_grad_(True)
        out, _ = blk(xd, ea, ei, ([[nd, D]], [[nd, D]], [[ea.shape[0], ed]]), 1)
        ref_out, grads = t(z["nodes_new"]), [(xd, "dx"), (ea, "dea")]
    (out * t(z["gd"])).sum().backward()
    assert torch.allclose(out, ref_out, atol=2e-6)
    for x, key in grads:
        assert torch.allclose(x.grad, t(z[key]), atol=5e-6), key
    for name, p in blk.named_parameters():
        ref = t(z["gp." + name])
        assert p.grad is not None and torch.allclose(p.grad, ref, atol=2e-5 * max(1.0, float(ref.abs().max()))), name


def test_wgrad_split_choice_fills_whole_waves():
    """Split-K of the weight-gradient GEMM: every split non-empty with >= 8 k-blocks, and the tile count (output tiles x
    splits) fills whole waves of the 74 CTA pairs at the shapes of the blocks."""
    from anemoi_models_b200.gemm import wgrad_splits

    for N, K, M in ((1024, 1024, 40320), (2048, 1024, 542080), (4096, 1024, 40320), (1024, 4096, 40320), (512, 512, 327660),
                    (1024, 16, 748256), (256, 256, 100), (64, 64, 8), (512, 1536, 40962)):
        s = wgrad_splits(N, K, M)
        kb = (M + 63) // 64
        assert s >= 1 and (s == 1 or (kb // s >= 8 and ((kb + s - 1) // s) * (s - 1) < kb)), (N, K, M, s)
        pairs = ((N + 255) // 256) * ((K + 255) // 256 if K > 128 else 1)
        tiles = pairs * s
        if kb >= 8 * 74:  # enough contraction to choose freely: at least 85 % of the last wave is used
            assert tiles / (((tiles + 73) // 74) * 74) >= 0.85, (N, K, M, s, tiles)


def test_o1280_recipe_bands_partition_the_whole_graph_and_balanced_bounds_balance():
    """The per-rank band generator of the o1280 -> n320 benchmark graph (here a small analogue: o24 -> 3,000 Fibonacci points):
    the bands of P = 1, 2, 5 partition the same edge set, edges are grouped by dst and sorted by src inside a dst, and the
    edge-balanced cut points give every rank the same edge count to within one dst row's degree."""
    import numpy as np

    from anemoi_models_b200 import synthetic as S

    nd = 3000
    radius = 0.6 * S.fibonacci_max_nn_distance(nd)
    whole, ns, nd_, _ = S.o1280_to_n320_band(1, 0, src_N=24, dst_points=nd, radius=radius)
    assert (ns, nd_) == (S.octahedral_size(24), nd) and whole.shape[1] > nd
    assert np.all(np.diff(whole[1]) >= 0)
    same_dst = np.diff(whole[1]) == 0
    assert np.all(np.diff(whole[0])[same_dst] > 0)
    for P in (2, 5):
        parts = [S.o1280_to_n320_band(P, r, src_N=24, dst_points=nd, radius=radius)[0] for r in range(P)]
        assert np.array_equal(np.concatenate(parts, axis=1), whole)
    deg = np.bincount(whole[1], minlength=nd)
    b = S.edge_balanced_bounds(deg, 4)
    assert b[0] == 0 and b[-1] == nd and all(x <= y for x, y in zip(b[:-1], b[1:]))
    counts = [int(deg[b[i]:b[i + 1]].sum()) for i in range(4)]
    assert max(counts) - min(counts) <= 2 * int(deg.max())
    bal = [S.o1280_to_n320_band(4, r, src_N=24, dst_points=nd, radius=radius, dst_bounds=b)[0] for r in range(4)]
    assert np.array_equal(np.concatenate(bal, axis=1), whole)


def test_encoder_work_balanced_bounds_even_out_the_weak_scaling_shards():
    """The weak-scaling headline shards the `fibonacci -> oN` encoder graph by dst rows.  Equal-count dst shards of an octahedral
    grid are not equal-area, so the polar ranks hold more src rows than the single-GPU workload; the closed-form work-balanced
    cut points (no graph needed; measured weights: an edge costs 1.7 src rows) bring the busiest rank's work to within 2 % of the
    mean, sit on latitude-row boundaries, and the bands they define still partition the whole edge set."""
    import numpy as np

    from anemoi_models_b200 import synthetic as S

    P, N, ns = 8, 40, 8 * 20000
    whole = S.encoder_graph_band(ns, N, 1, 0)[0]
    b = S.encoder_work_balanced_bounds(ns, N, P)
    assert b[0] == 0 and b[-1] == S.octahedral_size(N) and all(x < y for x, y in zip(b[:-1], b[1:]))
    row_starts = set(np.concatenate([[0], np.cumsum(S.octahedral_rows(N)[1])]).tolist())
    assert all(c in row_starts for c in b)
    assert S.encoder_work_balanced_bounds(2 * 20000, N, 2)[1] == S.octahedral_size(N) // 2  # two ranks: the equator, as tensor_split

    def work(bounds):
        bands = [S.encoder_graph_band(ns, N, P, r, bounds=bounds)[0] for r in range(P)]
        assert np.array_equal(np.concatenate(bands, axis=1), whole)
        w = np.array([1.7 * e.shape[1] + (e[0].max() - e[0].min() + 1) for e in bands])
        return float(w.max() / w.mean())

    equal, balanced = work(None), work(b)
    assert equal > 1.03 and balanced < 1.02 and balanced < equal, (equal, balanced)


def _emulated_gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, lda=None, ldb=None, out=None, seg_cols=0, out_dtype=torch.bfloat16,
                   bias=None, row_scale=None, row_shift=None, col_vec=None, act=3, pre_out=None, dact_pre=None, residual=None,
                   splits=1, gather=None, a_segs=None, a_seg_len=0):
    """torch restatement of what `ab2_gemm_bf16` computes (include/anemoi_b200.h), operand layouts and epilogue order included"""
    F = torch.nn.functional
    fn = {0: F.silu, 1: F.gelu, 2: F.relu}
    if a_segs:  # segmented A: pieces of a_seg_len along K (K-major) or along M (MN-major) -- the inner dimension either way
        assert all(t.shape == a.shape and t.shape[1] == a_seg_len for t in a_segs)
        a = torch.cat([a, *a_segs], dim=1)
    A = (a.float().t() if a_mn else a.float())[:M, :K]
    B = (b.float().t() if b_mn else b.float())[:N, :K]
    acc = A @ B.t()
    if row_scale is not None:
        acc = row_scale[:, None] * acc + row_shift[:, None] * col_vec[None, :]
    if bias is not None:
        acc = acc + bias
    if gather is not None:
        (ta, ia), (tb, ib) = gather
        acc = acc + ta.float()[ia] + tb.float()[ib]
    if dact_pre is not None:
        x = dact_pre.float().requires_grad_(True)
        with torch.enable_grad():
            (grad,) = torch.autograd.grad(fn[act](x).sum(), x)
        acc = acc * grad
    elif act != 3:
        if pre_out is not None:
            pre_out.copy_(acc.to(torch.bfloat16))
            acc = pre_out.float()
        acc = fn[act](acc)
    if residual is not None:
        acc = acc + residual.float()
    res = acc.to(out_dtype)
    if out is not None:
        for i, o in enumerate(out):
            o.copy_(res[:, i * seg_cols:(i + 1) * seg_cols])
        return out
    return res


def test_tensor_core_autograd_glue_against_torch_on_the_cpu(monkeypatch):
    """The autograd functions of gemm.py (which operand goes where in forward, dgrad and wgrad; which tensors are kept; how the
    activation derivative, the bias gradient and the GraphConv node-table gradients are assembled) with the three device entry
    points replaced by torch restatements: outputs and every gradient against plain fp32 autograd on bf16-rounded parameters."""
    from anemoi_models_b200 import gemm as G

    F = torch.nn.functional
    monkeypatch.setattr(G, "gemm", _emulated_gemm)
    monkeypatch.setattr(G, "colsum", lambda a_: a_.float().sum(0))

    def ln_fwd(x2, gamma, beta, eps, out_dtype):
        mean, var = x2.float().mean(1), x2.float().var(1, unbiased=False)
        rstd = (var + eps).rsqrt()
        return (((x2.float() - mean[:, None]) * rstd[:, None]) * gamma + beta).to(out_dtype), mean, rstd

    def ln_bwd(g2, x2, gamma, mean, rstd, add=None):
        xh = (x2.float() - mean[:, None]) * rstd[:, None]
        gg = g2.float() * gamma
        dx = rstd[:, None] * (gg - gg.mean(1, keepdim=True) - xh * (gg * xh).mean(1, keepdim=True))
        if add is not None:
            assert add.dtype == x2.dtype and add.shape == x2.shape
            dx = dx + add.float()
        return dx.to(x2.dtype), (g2.float() * xh).sum(0), g2.float().sum(0)

    def seg(g2, plan, want_dst, want_src):
        ei = plan.edge_index
        dpi = torch.zeros(plan.num_dst, g2.shape[1]).index_add_(0, ei[1], g2.float()).bfloat16() if want_dst else None
        dpj = torch.zeros(plan.num_src, g2.shape[1]).index_add_(0, ei[0], g2.float()).bfloat16() if want_src else None
        return dpi, dpj

    monkeypatch.setattr(G, "_ln_fwd_kernel", ln_fwd)
    monkeypatch.setattr(G, "_ln_bwd_kernel", ln_bwd)
    monkeypatch.setattr(G, "_segment_sums_kernel", seg)
    torch.manual_seed(0)
    M, D, Hd = 60, 16, 32
    ln, l1, l2 = torch.nn.LayerNorm(D), torch.nn.Linear(D, Hd), torch.nn.Linear(Hd, D)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
    x, g = torch.randn(M, D), torch.randn(M, D)

    def close(a_, b_, what):
        err = float((a_.float() - b_.float()).abs().max() / max(1.0, float(b_.float().abs().max())))
        assert err < 2e-2, (what, err)

    # LayerNorm -> Linear + GELU (epilogue) -> Linear + residual: the node MLP of the GT blocks
    for name, fn in (("GELU", F.gelu), ("SiLU", F.silu)):
        act = G.ACT_CODES[name]
        for m in (ln, l1, l2):
            m.zero_grad()
        x1 = x.clone().requires_grad_(True)
        pre, h = G.linear(G.layer_norm(x1, ln), l1, act_out=act)
        y = G.act_linear(pre, h, l2, act, residual=x1)
        y.backward(g)
        got = [y, x1.grad, ln.weight.grad.clone(), ln.bias.grad.clone(), l1.weight.grad.clone(), l1.bias.grad.clone(),
               l2.weight.grad.clone(), l2.bias.grad.clone()]
        for m in (ln, l1, l2):
            m.zero_grad()
        xr = x.clone().requires_grad_(True)
        yr = l2(fn(l1(ln(xr)))) + xr
        yr.backward(g)
        ref = [yr, xr.grad, ln.weight.grad, ln.bias.grad, l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad]
        for i, (a_, b_) in enumerate(zip(got, ref)):
            close(a_, b_, (name, i))
    # the GT block's pre-norm attention input: (LN(x), skip) fork + q | k | v | self as ONE segmented GEMM each way
    Dm = 256  # linear_multi fuses layers whose width is a multiple of 256
    lnm = torch.nn.LayerNorm(Dm)
    lins = [torch.nn.Linear(Dm, Dm) for _ in range(4)]
    xm, gm = torch.randn(M, Dm), [torch.randn(M, Dm) for _ in range(5)]

    def run(fused):
        for m in [lnm] + lins:
            m.zero_grad()
        xi = xm.clone().requires_grad_(True)
        if fused:
            xn, skip = G.layer_norm_fork(xi, lnm)
            ys = G.linear_multi(xn, lins)
            assert len(ys) == 4 and G._MultiLinearFn.__name__ in type(ys[0].grad_fn).__name__
        else:
            xn, skip = lnm(xi), xi
            ys = [l(xn) for l in lins]
        torch.autograd.backward(list(ys) + [skip], gm)
        return list(ys) + [xi.grad] + [p.grad.clone() for m in [lnm] + lins for p in m.parameters()]

    for i, (a_, b_) in enumerate(zip(run(True), run(False))):
        close(a_, b_, ("linear_multi", i))
    two = G.linear_multi(G.layer_norm(xm, lnm), lins[:2])  # the mapper's pairs (k | v on src, q | self on dst)
    close(two[1], lins[1](lnm(xm)), "linear_multi pair")
    odd = [torch.nn.Linear(D, Hd) for _ in range(2)]  # widths the segmented kernel does not take: one GEMM per layer
    assert all("_LinearFn" in type(y.grad_fn).__name__ for y in G.linear_multi(x, odd))

    # GraphConv first layer on the split weight: pre = e We^T + pi[dst] + pj[src]
    ns, nd, E = 9, 7, 40
    ei = torch.stack([torch.randint(0, ns, (E,)), torch.randint(0, nd, (E,))])

    class Plan:
        edge_index, num_src, num_dst = ei, ns, nd

    W = torch.randn(Hd, 3 * D, requires_grad=True)
    e, xs, xd = (torch.randn(n, D, requires_grad=True) for n in (E, ns, nd))
    pi, pj = F.linear(xd, W[:, :D]), F.linear(xs, W[:, D:2 * D])
    pre, h = G.edge_first_layer(e, pi, pj, W[:, 2 * D:], Plan, G.ACT_CODES["SiLU"])
    out = G.act_linear(pre, h, l2, G.ACT_CODES["SiLU"])
    ge = torch.randn(E, D)
    out.backward(ge)
    got = [out, e.grad.clone(), xs.grad.clone(), xd.grad.clone(), W.grad.clone()]
    for t_ in (e, xs, xd, W):
        t_.grad = None
    ref_out = l2(F.silu(F.linear(torch.cat([xd[ei[1]], xs[ei[0]], e], 1), W)))
    ref_out.backward(ge)
    for i, (a_, b_) in enumerate(zip(got, [ref_out, e.grad, xs.grad, xd.grad, W.grad])):
        close(a_, b_, ("edge_first_layer", i))
