"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

The reference's Python files are imported unchanged from /root/reference/src; torch-geometric (absent
from the image) is provided by the restated shim in oracle/pyg_shim.  The reference's own tests hold
no golden vectors for this path, so these fixtures -- outputs of the reference itself on
seeded inputs -- are what pins the oracle and the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
sys.path[:0] = [os.path.join(HERE, "pyg_shim"), REF, ROOT]

from anemoi.models.distributed.khop_edges import sort_edges_1hop_chunks  # noqa: E402
from anemoi.models.distributed.shapes import get_shape_shards  # noqa: E402
from anemoi.models.layers.block import (  # noqa: E402
    GraphConvMapperBlock,
    GraphConvProcessorBlock,
    GraphTransformerMapperBlock,
    GraphTransformerProcessorBlock,
)
from anemoi.models.layers.conv import GraphConv, GraphTransformerConv  # noqa: E402

from oracle import blocks as oblocks  # noqa: E402
from oracle import gtconv as og  # noqa: E402
from oracle import sharding as osh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rand_graph(gen, ns, nd, e, isolated=(), dup=0):
    src = torch.randint(0, ns, (e,), generator=gen)
    allowed = torch.tensor([i for i in range(nd) if i not in set(isolated)])
    dst = allowed[torch.randint(0, len(allowed), (e,), generator=gen)]
    ei = torch.stack([src, dst])
    if dup:
        ei = torch.cat([ei, ei[:, :dup]], dim=1)  # exact duplicate edges
    return ei.to(torch.int64)


def npz(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        conv[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name), **conv)
    print(f"wrote {name}: " + ", ".join(f"{k}{list(np.shape(v))}" for k, v in conv.items()))


def ref_gtconv(q, k, v, e, ei, g, size):
    q, k, v, e = (t.clone().requires_grad_(True) for t in (q, k, v, e))
    C = q.shape[2]
    out = GraphTransformerConv(out_channels=C)(q, k, v, e, ei, size)
    out.backward(g)
    return out.detach(), q.grad, k.grad, v.grad, e.grad


def case_gtconv(name, seed, ns, nd, E, H, C, isolated=(), dup=0, scale=1.0, size=True):
    gen = torch.Generator().manual_seed(seed)
    ei = rand_graph(gen, ns, nd, E, isolated, dup)
    Et = ei.shape[1]
    q = torch.randn(nd, H, C, generator=gen) * scale
    k = torch.randn(ns, H, C, generator=gen) * scale
    v = torch.randn(ns, H, C, generator=gen)
    e = torch.randn(Et, H, C, generator=gen)
    g = torch.randn(nd, H, C, generator=gen)
    out, dq, dk, dv, de = ref_gtconv(q, k, v, e, ei, g, (ns, nd) if size else None)
    # pin the oracle restatements against the reference right here
    o1 = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns, nd))
    for a, b in ((o1["out"], out), (o1["dq"], dq), (o1["dk"], dk), (o1["dv"], dv), (o1["de"], de)):
        assert torch.equal(a, b), f"{name}: unfused oracle differs from the reference"
    o2 = og.gt_conv_csr_f64(q, k, v, e, ei, g)
    for key, b in (("out", out), ("dq", dq), ("dk", dk), ("dv", dv), ("de", de)):
        err = np.abs(o2[key] - b.double().numpy()).max() / max(1.0, float(b.abs().max()))
        assert err < 2e-5, f"{name}: csr f64 oracle vs reference {key}: {err}"
    o3 = og.gt_conv_loops_f64(q, k, v, e, ei)
    assert np.abs(o3 - o2["out"]).max() < 1e-10, f"{name}: loops oracle vs csr oracle"
    npz(name, q=q, k=k, v=v, e=e, edge_index=ei, g=g, out=out, dq=dq, dk=dk, dv=dv, de=de, size=np.array([ns, nd]))


def kat_three_nodes():
    """Hand-computable: 2 src, 3 dst (dst 2 isolated), H=1, C=2.  dst0 has ONE incoming edge (alpha=1 exactly);
    dst1 has two edges with equal logits (alpha = 1/2 each)."""
    q = torch.tensor([[[1.0, 0.0]], [[0.0, 2.0]], [[5.0, 5.0]]])
    k = torch.tensor([[[1.0, 1.0]], [[3.0, 1.0]]])
    v = torch.tensor([[[1.0, 2.0]], [[3.0, 4.0]]])
    ei = torch.tensor([[0, 0, 1], [0, 1, 1]])
    e = torch.tensor([[[0.5, -0.5]], [[0.0, 0.0]], [[1.0, 0.0]]])
    g = torch.ones(3, 1, 2)
    out, dq, dk, dv, de = ref_gtconv(q, k, v, e, ei, g, (2, 3))
    expect = torch.tensor([[[1.5, 1.5]], [[2.5, 3.0]], [[0.0, 0.0]]])  # by hand
    assert torch.allclose(out, expect, atol=1e-6), out
    npz("gtconv_kat3.npz", q=q, k=k, v=v, e=e, edge_index=ei, g=g, out=out, dq=dq, dk=dk, dv=dv, de=de,
        size=np.array([2, 3]), expect=expect)


def state_np(mod):
    return {k: v.detach().clone() for k, v in mod.state_dict().items()}


def case_graphconv(seed=5):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    ns, nd, E, D = 30, 20, 90, 16
    ei = rand_graph(gen, ns, nd, E, isolated=(3,), dup=4)
    conv = GraphConv(in_channels=D, out_channels=D)
    xs = torch.randn(ns, D, generator=gen, requires_grad=True)
    xd = torch.randn(nd, D, generator=gen, requires_grad=True)
    e = torch.randn(ei.shape[1], D, generator=gen, requires_grad=True)
    go = torch.randn(nd, D, generator=gen)
    ge = torch.randn(ei.shape[1], D, generator=gen)
    out, en = conv((xs, xd), e, ei, size=(ns, nd))
    (out * go).sum().add((en * ge).sum()).backward()
    p = state_np(conv)
    o_out, o_en = og.graph_conv_unfused((xs.detach(), xd.detach()), e.detach(), ei, p, "edge_mlp.")
    assert torch.equal(o_out, out.detach()) and torch.equal(o_en, en.detach()), "graph_conv oracle differs from reference"
    arrs = {f"p.{k}": v for k, v in p.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in conv.named_parameters()})
    npz("graphconv_bipartite.npz", xs=xs, xd=xd, e=e, edge_index=ei, go=go, ge=ge, out=out, edges_new=en,
        dxs=xs.grad, dxd=xd.grad, de=e.grad, size=np.array([ns, nd]), **arrs)
    # single node set (processor use)
    torch.manual_seed(seed + 1)
    n = 25
    ei2 = rand_graph(gen, n, n, 80)
    conv2 = GraphConv(in_channels=D, out_channels=D)
    x = torch.randn(n, D, generator=gen, requires_grad=True)
    e2 = torch.randn(80, D, generator=gen, requires_grad=True)
    out2, en2 = conv2(x, e2, ei2)
    (out2.square().sum() + en2.square().sum()).backward()
    p2 = state_np(conv2)
    arrs = {f"p.{k}": v for k, v in p2.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in conv2.named_parameters()})
    npz("graphconv_single.npz", x=x, e=e2, edge_index=ei2, out=out2, edges_new=en2, dx=x.grad, de=e2.grad, **arrs)


def case_blocks(seed=11):
    gen = torch.Generator().manual_seed(seed)
    # --- GT mapper block
    torch.manual_seed(seed)
    ns, nd, E, D, H, ed, hid = 36, 20, 110, 32, 4, 5, 64
    ei = rand_graph(gen, ns, nd, E, isolated=(7,), dup=3)
    blk = GraphTransformerMapperBlock(in_channels=D, hidden_dim=hid, out_channels=D, edge_dim=ed, num_heads=H)
    xs = torch.randn(ns, D, generator=gen, requires_grad=True)
    xd = torch.randn(nd, D, generator=gen, requires_grad=True)
    ea = torch.rand(ei.shape[1], ed, generator=gen, requires_grad=True)
    shapes = ([[ns, D]], [[nd, D]], [[ei.shape[1], ed]])
    (src_new, dst_new), ea_out = blk((xs, xd), ea, ei, shapes, 1, size=(ns, nd))
    gd = torch.randn(nd, D, generator=gen)
    (dst_new * gd).sum().backward()
    p = state_np(blk)
    (o_src, o_dst), _ = oblocks.gt_mapper_block(p, (xs.detach(), xd.detach()), ea.detach(), ei, H, (ns, nd))
    assert torch.equal(o_dst, dst_new.detach()), "gt mapper block oracle differs from the reference"
    arrs = {f"p.{k}": v for k, v in p.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in blk.named_parameters()})
    npz("block_gt_mapper.npz", xs=xs, xd=xd, ea=ea, edge_index=ei, gd=gd, dst_new=dst_new, dxs=xs.grad, dxd=xd.grad,
        dea=ea.grad, meta=np.array([ns, nd, D, H, ed, hid]), **arrs)
    # --- GT processor block
    torch.manual_seed(seed + 1)
    n = 28
    ei = rand_graph(gen, n, n, 100, dup=2)
    blk = GraphTransformerProcessorBlock(in_channels=D, hidden_dim=hid, out_channels=D, edge_dim=ed, num_heads=H)
    x = torch.randn(n, D, generator=gen, requires_grad=True)
    ea = torch.rand(ei.shape[1], ed, generator=gen, requires_grad=True)
    shapes = ([[n, D]], [[n, D]], [[ei.shape[1], ed]])
    nodes_new, _ = blk(x, ea, ei, shapes, 1)
    gd = torch.randn(n, D, generator=gen)
    (nodes_new * gd).sum().backward()
    p = state_np(blk)
    o_nodes, _ = oblocks.gt_processor_block(p, x.detach(), ea.detach(), ei, H)
    assert torch.equal(o_nodes, nodes_new.detach()), "gt processor block oracle differs from the reference"
    arrs = {f"p.{k}": v for k, v in p.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in blk.named_parameters()})
    npz("block_gt_processor.npz", x=x, ea=ea, edge_index=ei, gd=gd, nodes_new=nodes_new, dx=x.grad, dea=ea.grad,
        meta=np.array([n, n, D, H, ed, hid]), **arrs)
    # --- GraphConv processor block
    torch.manual_seed(seed + 2)
    ei = rand_graph(gen, n, n, 90)
    blk = GraphConvProcessorBlock(in_channels=D, out_channels=D)
    x = torch.randn(n, D, generator=gen, requires_grad=True)
    e = torch.randn(90, D, generator=gen, requires_grad=True)
    nodes_new, edges_new = blk(x, e, ei, ([[n, D]], [[n, D]], [[90, D]]))
    gd = torch.randn(n, D, generator=gen)
    ge = torch.randn(90, D, generator=gen)
    ((nodes_new * gd).sum() + (edges_new * ge).sum()).backward()
    p = state_np(blk)
    o_nodes, o_edges = oblocks.graphconv_processor_block(p, x.detach(), e.detach(), ei)
    assert torch.equal(o_nodes, nodes_new.detach()) and torch.equal(o_edges, edges_new.detach())
    arrs = {f"p.{k}": v for k, v in p.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in blk.named_parameters()})
    npz("block_graphconv_processor.npz", x=x, e=e, edge_index=ei, gd=gd, ge=ge, nodes_new=nodes_new, edges_new=edges_new,
        dx=x.grad, de=e.grad, meta=np.array([n, D]), **arrs)
    # --- GraphConv mapper block
    torch.manual_seed(seed + 3)
    ei = rand_graph(gen, ns, nd, 100, isolated=(2,))
    blk = GraphConvMapperBlock(in_channels=D, out_channels=D)
    xs = torch.randn(ns, D, generator=gen, requires_grad=True)
    xd = torch.randn(nd, D, generator=gen, requires_grad=True)
    e = torch.randn(100, D, generator=gen, requires_grad=True)
    (src_new, dst_new), edges_new = blk((xs, xd), e, ei, ([[ns, D]], [[nd, D]], [[100, D]]), size=(ns, nd))
    gs = torch.randn(ns, D, generator=gen)
    gd = torch.randn(nd, D, generator=gen)
    ((src_new * gs).sum() + (dst_new * gd).sum()).backward()
    p = state_np(blk)
    (o_s, o_d), o_e = oblocks.graphconv_mapper_block(p, (xs.detach(), xd.detach()), e.detach(), ei, (ns, nd))
    assert torch.equal(o_s, src_new.detach()) and torch.equal(o_d, dst_new.detach())
    arrs = {f"p.{k}": v for k, v in p.items()}
    arrs.update({f"gp.{k}": v.grad for k, v in blk.named_parameters()})
    npz("block_graphconv_mapper.npz", xs=xs, xd=xd, e=e, edge_index=ei, gs=gs, gd=gd, src_new=src_new, dst_new=dst_new,
        edges_new=edges_new, dxs=xs.grad, dxd=xd.grad, de=e.grad, meta=np.array([ns, nd, D]), **arrs)


def case_sharding(seed=21):
    gen = torch.Generator().manual_seed(seed)
    arrs = {}
    # get_shape_shards without a group = 1 shard; the sharded shapes are torch.tensor_split shapes (shapes.py:24)
    for n, P in ((10, 3), (40320, 8), (7, 8), (542080, 8)):
        ref = [list(x.shape) for x in torch.tensor_split(torch.empty(n, 1), P, dim=0)]
        assert ref == osh.shape_shards((n, 1), 0, P)
        arrs[f"shards_{n}_{P}"] = np.array(ref)
    assert get_shape_shards(torch.empty(9, 4), 0, None) == [[9, 4]]
    # sort_edges_1hop_chunks, bipartite (tuple) and single node set (int)
    ns, nd, E = 50, 23, 300
    ei = rand_graph(gen, ns, nd, E, isolated=(0, 11), dup=5)
    ea = torch.arange(ei.shape[1]).view(-1, 1).float()  # the attr IS the original edge id
    arrs["bip_edge_index"] = ei.numpy()
    for P in (1, 2, 3, 4, 8):
        ea_list, ei_list = sort_edges_1hop_chunks((ns, nd), ea, ei, P)
        ids = [a.view(-1).long().numpy() for a in ea_list]
        mine = osh.edges_1hop_chunks((ns, nd), ei.numpy(), P)
        for c in range(P):
            assert np.array_equal(ids[c], mine[c]) and np.array_equal(ei_list[c].numpy(), ei.numpy()[:, mine[c]])
        arrs[f"bip_ids_P{P}"] = np.concatenate(ids)
        arrs[f"bip_counts_P{P}"] = np.array([len(i) for i in ids])
    n = 31
    ei = rand_graph(gen, n, n, 200)
    ei[1, 0] = n - 1  # make sure edge_index.max()+1 == n (PyG k_hop_subgraph infers num_nodes from it)
    ea = torch.arange(ei.shape[1]).view(-1, 1).float()
    arrs["one_edge_index"] = ei.numpy()
    for P in (1, 2, 4, 5):
        ea_list, ei_list = sort_edges_1hop_chunks(n, ea, ei, P)
        ids = [a.view(-1).long().numpy() for a in ea_list]
        mine = osh.edges_1hop_chunks(n, ei.numpy(), P)
        for c in range(P):
            assert np.array_equal(ids[c], mine[c]) and np.array_equal(ei_list[c].numpy(), ei.numpy()[:, mine[c]])
        arrs[f"one_ids_P{P}"] = np.concatenate(ids)
        arrs[f"one_counts_P{P}"] = np.array([len(i) for i in ids])
    arrs["meta"] = np.array([ns, nd, n])
    npz("sharding.npz", **arrs)


def main():
    torch.set_num_threads(1)  # deterministic reductions in the fixtures
    case_gtconv("gtconv_bipartite.npz", 1, ns=40, nd=24, E=150, H=4, C=8, isolated=(5, 17), dup=6)
    case_gtconv("gtconv_c64.npz", 2, ns=20, nd=12, E=70, H=2, C=64, isolated=(0,), dup=2)
    case_gtconv("gtconv_c16_h16.npz", 3, ns=24, nd=16, E=100, H=16, C=16, isolated=(15,), size=False)
    case_gtconv("gtconv_biglogit.npz", 4, ns=16, nd=8, E=64, H=2, C=8, scale=6.0)  # logits O(100): max-subtraction matters
    case_gtconv("gtconv_oddc.npz", 6, ns=14, nd=9, E=40, H=3, C=5)  # C not a multiple of the vector width
    kat_three_nodes()
    case_graphconv()
    case_blocks()
    case_sharding()


if __name__ == "__main__":
    main()
