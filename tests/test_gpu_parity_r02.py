"""Parity cases the round-1 suite left open ("what's weak" 1a / 1b):
  * the bf16 instantiations of the GraphConv kernels and both GraphConv blocks against the fp32 oracle evaluated on the same
    bf16-rounded inputs and parameters (D = 512, multi-scale icosahedral mesh as in BASELINE configs[2], refinement 4);
  * the ACTUAL headline graph (synthetic.encoder_graph(542080, 96): E = 748,256, D = 1024, H = 16, bf16) against the reference's
    unfused op sequence in fp32 executed on the same GPU.
Tolerance for bf16: 2e-2 of max|ref| (the measure of the round-1 tests) AND 2e-2 relative L2 on outputs and every gradient."""
import pytest
import torch

from conftest import rel_err, rel_l2
from oracle import blocks as oblocks
from oracle import gtconv as og

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _mesh(refinement=4):
    from anemoi_models_b200 import synthetic as S

    xyz, ei = S.multiscale_icosahedral_mesh(refinement)
    return len(xyz), torch.from_numpy(ei)


def _bf16_round_state(mod):
    """module parameters rounded to bf16 (as fp32 tensors, requires_grad) keyed like the state_dict"""
    return {k: v.detach().bfloat16().float().cpu().requires_grad_(True) for k, v in mod.state_dict().items()}


def _check(named_pairs):
    for name, got, ref in named_pairs:
        e1, e2 = rel_err(got.float(), ref), rel_l2(got.float(), ref)
        assert e1 < TOL and e2 < TOL, (name, e1, e2)


@pytest.mark.parametrize("act,extra", [("SiLU", 0), ("GELU", 1)])
def test_graphconv_bf16_against_fp32_oracle(act, extra):
    import anemoi_models_b200 as b2

    torch.manual_seed(0)
    n, ei = _mesh(4)
    D, E = 512, ei.shape[1]
    conv = b2.GraphConv(D, D, mlp_extra_layers=extra, activation=act).cuda().bfloat16()
    p = _bf16_round_state(conv)
    gen = torch.Generator().manual_seed(1)
    x, e = torch.randn(n, D, generator=gen).bfloat16(), torch.randn(E, D, generator=gen).bfloat16()
    go, ge = torch.randn(n, D, generator=gen).bfloat16(), torch.randn(E, D, generator=gen).bfloat16()
    xr, er = x.float().requires_grad_(True), e.float().requires_grad_(True)
    o_ref, en_ref = og.graph_conv_unfused(xr, er, ei, p, "edge_mlp.", extra, act, (n, n))
    ((o_ref * go.float()).sum() + (en_ref * ge.float()).sum()).backward()
    xc, ec = x.cuda().requires_grad_(True), e.cuda().requires_grad_(True)
    out, en = conv(xc, ec, ei.cuda(), (n, n))
    assert out.dtype == torch.bfloat16 and en.dtype == torch.bfloat16
    ((out * go.cuda()).sum() + (en * ge.cuda()).sum()).backward()
    pairs = [("out", out, o_ref), ("edges_new", en, en_ref), ("dx", xc.grad, xr.grad), ("de", ec.grad, er.grad)]
    pairs += [(k, prm.grad, p[k].grad) for k, prm in conv.named_parameters()]
    _check(pairs)


def test_graphconv_blocks_bf16_against_fp32_oracle():
    import anemoi_models_b200 as b2

    torch.manual_seed(2)
    n, ei = _mesh(4)
    D, E = 512, ei.shape[1]
    gen = torch.Generator().manual_seed(3)
    # ---- processor block (one node set; edge features are updated and carried to the next layer, block.py:223)
    blk = b2.GraphConvProcessorBlock(D, D).cuda().bfloat16()
    p = _bf16_round_state(blk)
    x, e = torch.randn(n, D, generator=gen).bfloat16(), torch.randn(E, D, generator=gen).bfloat16()
    gn, ge = torch.randn(n, D, generator=gen).bfloat16(), torch.randn(E, D, generator=gen).bfloat16()
    xr, er = x.float().requires_grad_(True), e.float().requires_grad_(True)
    n_ref, e_ref = oblocks.graphconv_processor_block(p, xr, er, ei)
    ((n_ref * gn.float()).sum() + (e_ref * ge.float()).sum()).backward()
    xc, ec = x.cuda().requires_grad_(True), e.cuda().requires_grad_(True)
    nodes, edges = blk(xc, ec, ei.cuda(), ([[n, D]], [[n, D]], [[E, D]]))
    ((nodes * gn.cuda()).sum() + (edges * ge.cuda()).sum()).backward()
    pairs = [("nodes_new", nodes, n_ref), ("edges_new", edges, e_ref), ("dx", xc.grad, xr.grad), ("de", ec.grad, er.grad)]
    pairs += [(k, prm.grad, p[k].grad) for k, prm in blk.named_parameters()]
    _check(pairs)
    # ---- mapper block (bipartite: coarse nodes of the mesh -> all nodes), update_src_nodes default True
    ns = 642  # nodes of refinement 3 come first in the multi-scale numbering
    keep = ei[0] < ns
    eb = ei[:, keep]
    Eb = eb.shape[1]
    blk = b2.GraphConvMapperBlock(D, D).cuda().bfloat16()
    p = _bf16_round_state(blk)
    xs, xd, e = (torch.randn(m, D, generator=gen).bfloat16() for m in (ns, n, Eb))
    gs, gd, ge = (torch.randn(m, D, generator=gen).bfloat16() for m in (ns, n, Eb))
    xsr, xdr, er = (t_.float().requires_grad_(True) for t_ in (xs, xd, e))
    (s_ref, d_ref), e_ref = oblocks.graphconv_mapper_block(p, (xsr, xdr), er, eb, size=(ns, n))
    ((s_ref * gs.float()).sum() + (d_ref * gd.float()).sum() + (e_ref * ge.float()).sum()).backward()
    xsc, xdc, ec = (t_.cuda().requires_grad_(True) for t_ in (xs, xd, e))
    (s_new, d_new), e_new = blk((xsc, xdc), ec, eb.cuda(), ([[ns, D]], [[n, D]], [[Eb, D]]), size=(ns, n))
    ((s_new * gs.cuda()).sum() + (d_new * gd.cuda()).sum() + (e_new * ge.cuda()).sum()).backward()
    pairs = [("src_new", s_new, s_ref), ("dst_new", d_new, d_ref), ("edges_new", e_new, e_ref), ("dxs", xsc.grad, xsr.grad),
             ("dxd", xdc.grad, xdr.grad), ("de", ec.grad, er.grad)]
    pairs += [(k, prm.grad, p[k].grad) for k, prm in blk.named_parameters()]
    _check(pairs)


def test_headline_graph_bf16_against_unfused_fp32_on_gpu():
    """BASELINE configs[1] itself: Fibonacci(542,080) -> o96, cut-off 0.6, E = 748,256, D = 1024, H = 16, bf16 fwd+bwd through the
    C ABI, against the reference's op sequence (oracle port of conv.py + PyG softmax / scatter, plain torch ops) in fp32 on the
    same bf16-rounded inputs, on the same GPU (~35 GB of fp32 temporaries)."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200 import synthetic as S

    ei_np, Ns, Nd, _ = S.encoder_graph(542080, 96)
    assert (Ns, Nd) == (542080, 40320) and ei_np.shape[1] == 748256
    ei = torch.from_numpy(ei_np).cuda()
    E, H, C = ei.shape[1], 16, 64
    torch.manual_seed(0)
    q, k, v, e, g = (torch.randn(m, H, C, device="cuda").bfloat16() for m in (Nd, Ns, Ns, E, Nd))
    ins = [t_.detach().requires_grad_(True) for t_ in (q, k, v, e)]
    out = b2.GraphTransformerConv(out_channels=C)(*ins, ei, (Ns, Nd))
    out.backward(g)
    ref_in = [t_.detach().float().requires_grad_(True) for t_ in (q, k, v, e)]
    ref = og.gt_conv_unfused(*ref_in, ei, (Ns, Nd))
    ref.backward(g.float())
    pairs = [("out", out, ref.detach())] + [("d" + nm, a.grad, b.grad) for nm, a, b in zip("qkve", ins, ref_in)]
    for name, got, want in pairs:
        # chunked over rows: the fp64 copies of a [E, D] tensor would not fit next to the autograd graph
        mx, num, den, bmax = 0.0, 0.0, 0.0, 0.0
        step = 1 << 16
        for i in range(0, got.shape[0], step):
            d = got[i:i + step].double() - want[i:i + step].double()
            mx, bmax = max(mx, float(d.abs().max())), max(bmax, float(want[i:i + step].abs().max()))
            num, den = num + float((d * d).sum()), den + float((want[i:i + step].double() ** 2).sum())
        assert mx / max(1.0, bmax) < TOL and (num / den) ** 0.5 < TOL, (name, mx / max(1.0, bmax), (num / den) ** 0.5)
