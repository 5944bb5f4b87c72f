"""GraphConv and the four graph blocks on the GPU against the golden vectors of the reference (fp32 1e-5 rel;
dense GEMMs run in full fp32 -- TF32 is disabled for the comparison) and in bf16 (2e-2 rel)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, t
from oracle import blocks as oblocks
from oracle import gtconv as og

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.fixture(autouse=True)
def _full_fp32_gemms():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


def _params(z, prefix="p."):
    return {k[len(prefix):]: t(v) for k, v in z.items() if k.startswith(prefix)}


def _check_param_grads(mod, z, tol, prefix="gp."):
    for name, p in mod.named_parameters():
        ref = t(z[prefix + name])
        assert p.grad is not None, name
        assert rel_err(p.grad.float(), ref) < tol, (name, rel_err(p.grad.float(), ref))


def test_graphconv_bipartite_golden():
    import anemoi_models_b200 as b2

    z = load_golden("graphconv_bipartite.npz")
    D = z["xs"].shape[1]
    conv = b2.GraphConv(D, D).cuda()
    conv.load_state_dict(_params(z))
    xs, xd, e = (t(z[k]).cuda().requires_grad_(True) for k in ("xs", "xd", "e"))
    out, en = conv((xs, xd), e, t(z["edge_index"]).cuda(), size=tuple(int(x) for x in z["size"]))
    ((out * t(z["go"]).cuda()).sum() + (en * t(z["ge"]).cuda()).sum()).backward()
    assert rel_err(out, t(z["out"])) < FP32_TOL and rel_err(en, t(z["edges_new"])) < FP32_TOL
    assert rel_err(xs.grad, t(z["dxs"])) < FP32_TOL and rel_err(xd.grad, t(z["dxd"])) < FP32_TOL
    assert rel_err(e.grad, t(z["de"])) < FP32_TOL
    _check_param_grads(conv, z, FP32_TOL)


def test_graphconv_single_nodeset_golden():
    import anemoi_models_b200 as b2

    z = load_golden("graphconv_single.npz")
    D = z["x"].shape[1]
    conv = b2.GraphConv(D, D).cuda()
    conv.load_state_dict(_params(z))
    x, e = (t(z[k]).cuda().requires_grad_(True) for k in ("x", "e"))
    out, en = conv(x, e, t(z["edge_index"]).cuda())
    (out.square().sum() + en.square().sum()).backward()
    assert rel_err(out, t(z["out"])) < FP32_TOL and rel_err(en, t(z["edges_new"])) < FP32_TOL
    assert rel_err(x.grad, t(z["dx"])) < FP32_TOL and rel_err(e.grad, t(z["de"])) < FP32_TOL
    _check_param_grads(conv, z, FP32_TOL)


@pytest.mark.parametrize("D,act,extra", [(64, "SiLU", 0), (128, "GELU", 1), (24, "ReLU", 0), (6, "SiLU", 0)])
def test_graphconv_random_vs_oracle(D, act, extra):
    import anemoi_models_b200 as b2

    torch.manual_seed(D)
    ns, nd, E = 120, 70, 900
    conv = b2.GraphConv(D, D, mlp_extra_layers=extra, activation=act).cuda()
    p = {k: v.detach().cpu() for k, v in conv.state_dict().items()}
    ei = torch.stack([torch.randint(0, ns, (E,)), torch.randint(0, nd, (E,))])
    xs, xd, e = torch.randn(ns, D), torch.randn(nd, D), torch.randn(E, D)
    xs_r, xd_r, e_r = (x.clone().requires_grad_(True) for x in (xs, xd, e))
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    o_ref, en_ref = og.graph_conv_unfused((xs_r, xd_r), e_r, ei, pr, "edge_mlp.", extra, act, (ns, nd))
    (o_ref.sin().sum() + en_ref.cos().sum()).backward()
    xs_c, xd_c, e_c = (x.cuda().requires_grad_(True) for x in (xs, xd, e))
    out, en = conv((xs_c, xd_c), e_c, ei.cuda(), (ns, nd))
    (out.sin().sum() + en.cos().sum()).backward()
    tol = 2e-5  # a chain of three fp32 GEMMs in a different summation order
    assert rel_err(out, o_ref) < tol and rel_err(en, en_ref) < tol
    assert rel_err(xs_c.grad, xs_r.grad) < tol and rel_err(xd_c.grad, xd_r.grad) < tol and rel_err(e_c.grad, e_r.grad) < tol
    for name, prm in conv.named_parameters():
        assert rel_err(prm.grad, pr[name].grad) < 5e-5, name


def test_gt_mapper_block_golden():
    import anemoi_models_b200 as b2

    z = load_golden("block_gt_mapper.npz")
    ns, nd, D, H, ed, hid = (int(x) for x in z["meta"])
    blk = b2.GraphTransformerMapperBlock(D, hid, D, edge_dim=ed, num_heads=H).cuda()
    blk.load_state_dict(_params(z))
    xs, xd, ea = (t(z[k]).cuda().requires_grad_(True) for k in ("xs", "xd", "ea"))
    shapes = ([[ns, D]], [[nd, D]], [[ea.shape[0], ed]])
    (src_new, dst_new), ea_out = blk((xs, xd), ea, t(z["edge_index"]).cuda(), shapes, 1, size=(ns, nd))
    (dst_new * t(z["gd"]).cuda()).sum().backward()
    assert src_new is xs and ea_out is ea
    assert rel_err(dst_new, t(z["dst_new"])) < FP32_TOL
    assert rel_err(xs.grad, t(z["dxs"])) < FP32_TOL and rel_err(xd.grad, t(z["dxd"])) < FP32_TOL
    assert rel_err(ea.grad, t(z["dea"])) < FP32_TOL
    _check_param_grads(blk, z, 2e-5)


def test_gt_processor_block_golden_and_bf16():
    import anemoi_models_b200 as b2

    z = load_golden("block_gt_processor.npz")
    n, _, D, H, ed, hid = (int(x) for x in z["meta"])
    blk = b2.GraphTransformerProcessorBlock(D, hid, D, edge_dim=ed, num_heads=H).cuda()
    blk.load_state_dict(_params(z))
    x, ea = (t(z[k]).cuda().requires_grad_(True) for k in ("x", "ea"))
    shapes = ([[n, D]], [[n, D]], [[ea.shape[0], ed]])
    nodes_new, _ = blk(x, ea, t(z["edge_index"]).cuda(), shapes, 1)
    (nodes_new * t(z["gd"]).cuda()).sum().backward()
    assert rel_err(nodes_new, t(z["nodes_new"])) < FP32_TOL
    assert rel_err(x.grad, t(z["dx"])) < FP32_TOL and rel_err(ea.grad, t(z["dea"])) < FP32_TOL
    _check_param_grads(blk, z, 2e-5)
    # bf16 autocast (what the trainer does): within 2e-2 of the fp32 reference
    with torch.autocast("cuda", dtype=torch.bfloat16):
        nodes_bf, _ = blk(x.detach(), ea.detach(), t(z["edge_index"]).cuda(), shapes, 1)
    assert rel_err(nodes_bf.float(), t(z["nodes_new"])) < BF16_TOL


def test_graphconv_blocks_golden():
    import anemoi_models_b200 as b2

    z = load_golden("block_graphconv_processor.npz")
    n, D = (int(x) for x in z["meta"])
    blk = b2.GraphConvProcessorBlock(D, D).cuda()
    blk.load_state_dict(_params(z))
    x, e = (t(z[k]).cuda().requires_grad_(True) for k in ("x", "e"))
    nodes_new, edges_new = blk(x, e, t(z["edge_index"]).cuda(), ([[n, D]], [[n, D]], [[e.shape[0], D]]))
    ((nodes_new * t(z["gd"]).cuda()).sum() + (edges_new * t(z["ge"]).cuda()).sum()).backward()
    assert rel_err(nodes_new, t(z["nodes_new"])) < FP32_TOL and rel_err(edges_new, t(z["edges_new"])) < FP32_TOL
    assert rel_err(x.grad, t(z["dx"])) < FP32_TOL and rel_err(e.grad, t(z["de"])) < FP32_TOL
    _check_param_grads(blk, z, 2e-5)

    z = load_golden("block_graphconv_mapper.npz")
    ns, nd, D = (int(x) for x in z["meta"])
    blk = b2.GraphConvMapperBlock(D, D).cuda()
    blk.load_state_dict(_params(z))
    xs, xd, e = (t(z[k]).cuda().requires_grad_(True) for k in ("xs", "xd", "e"))
    (src_new, dst_new), edges_new = blk((xs, xd), e, t(z["edge_index"]).cuda(), ([[ns, D]], [[nd, D]], [[e.shape[0], D]]),
                                        size=(ns, nd))
    ((src_new * t(z["gs"]).cuda()).sum() + (dst_new * t(z["gd"]).cuda()).sum()).backward()
    assert rel_err(src_new, t(z["src_new"])) < FP32_TOL and rel_err(dst_new, t(z["dst_new"])) < FP32_TOL
    assert rel_err(edges_new, t(z["edges_new"])) < FP32_TOL
    assert rel_err(xs.grad, t(z["dxs"])) < FP32_TOL and rel_err(xd.grad, t(z["dxd"])) < FP32_TOL
    assert rel_err(e.grad, t(z["de"])) < FP32_TOL
    _check_param_grads(blk, z, 2e-5)


def test_host_buffer_entry_points_match_device_path():
    """ab2_gtconv_fwd_bwd_host and its streamed variant (pinned host in/out, copies inside the call) == the
    device-resident path, bit for bit; dst-sorted and shuffled edge lists."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200 import ops
    from anemoi_models_b200.graph import get_csr

    torch.manual_seed(3)
    ns, nd, E = 500, 200, 4000
    ei = torch.stack([torch.randint(0, ns, (E,)), torch.randint(0, nd, (E,))])
    ei_sorted = ei[:, torch.sort(ei[1], stable=True).indices]
    cases = [(e_, c_, 4, 16, dt) for e_, c_ in ((ei, 16), (ei_sorted, 1), (ei_sorted, 7), (ei_sorted, 16), (ei_sorted, 500))
             for dt in (torch.float32, torch.bfloat16)]
    cases += [(ei_sorted, 7, 16, 64, torch.bfloat16), (ei_sorted, 16, 16, 32, torch.float32)]  # 2 KB rows: pipelined kernels
    for edges, chunks, H, C, dtype in cases:
        edges = edges.cuda()
        for _ in (0,):
            q, k, v, e, g = (torch.randn(s, H, C).to(dtype).pin_memory() for s in (nd, ns, ns, E, nd))
            plan = get_csr(edges, ns, nd)
            assert plan.perm_is_identity == (edges is not ei.cuda() and bool(torch.equal(edges.cpu(), ei_sorted)))
            out, dq, dk, dv, de = ops.gt_conv_host(q, k, v, e, g, plan, nchunks=chunks)
            qd, kd, vd, ed = (x.cuda().requires_grad_(True) for x in (q, k, v, e))
            o = b2.GraphTransformerConv(C)(qd, kd, vd, ed, edges, (ns, nd))
            o.backward(g.cuda())
            for a, b_ in ((out, o), (dq, qd.grad), (dk, kd.grad), (dv, vd.grad), (de, ed.grad)):
                assert torch.equal(a, b_.detach().cpu())


def test_host_stream_meta_covers_every_row_once():
    from anemoi_models_b200 import ops
    from anemoi_models_b200.graph import GraphCSR

    torch.manual_seed(4)
    ns, nd = 300, 90
    dst = torch.sort(torch.randint(0, nd, (1000,))).values
    src = (dst * 3 + torch.randint(0, 30, (1000,))).clamp_(max=ns - 1)
    plan = GraphCSR(torch.stack([src, dst]).cuda(), ns, nd)
    meta = ops.host_stream_meta(plan, 8)
    assert meta[0, 0] == 0 and meta[-1, 1] == nd and meta[0, 2] == 0 and meta[-1, 3] == 1000
    assert torch.equal(meta[1:, 0], meta[:-1, 1]) and torch.equal(meta[1:, 2], meta[:-1, 3])
    assert bool((meta[1:, 4] >= meta[:-1, 4]).all()) and bool((meta[1:, 5] >= meta[:-1, 5]).all())
    rowptr = plan.rowptr.cpu()
    for c in range(8):
        d0, d1, p0, p1, smax, fin = (int(x) for x in meta[c, :6])
        assert p0 == int(rowptr[d0]) and p1 == int(rowptr[d1])
        assert int(src[:p1].max()) == smax if p1 > 0 else smax == -1
        later = src[p1:]
        assert fin == (int(later.min()) if later.numel() else ns)
