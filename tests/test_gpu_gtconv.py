"""Fused GraphTransformerConv on the GPU (through the C ABI) against the golden vectors of the reference and the
oracle.  Tolerances (north_star): fp32 1e-5 relative, bf16 2e-2 relative, on outputs and all gradients."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, t
from oracle import gtconv as og

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2
GT_CASES = ["gtconv_bipartite.npz", "gtconv_c64.npz", "gtconv_c16_h16.npz", "gtconv_biglogit.npz", "gtconv_oddc.npz",
            "gtconv_kat3.npz"]


def run_b2(q, k, v, e, ei, g, size, dtype=torch.float32):
    import anemoi_models_b200 as b2

    dev = "cuda"
    q, k, v, e = (x.to(dev, dtype).requires_grad_(True) for x in (q, k, v, e))
    conv = b2.GraphTransformerConv(out_channels=q.shape[2])
    out = conv(q, k, v, e, ei.to(dev), size)
    out.backward(g.to(dev, dtype))
    return {"out": out.detach(), "dq": q.grad, "dk": k.grad, "dv": v.grad, "de": e.grad}


@pytest.mark.parametrize("name", GT_CASES)
def test_golden_fp32(name):
    z = load_golden(name)
    size = tuple(int(x) for x in z["size"])
    r = run_b2(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), t(z["g"]), size)
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], t(z[key])) < FP32_TOL, (name, key, rel_err(r[key], t(z[key])))
    if name == "gtconv_kat3.npz":
        assert torch.equal(r["out"][2].cpu(), torch.zeros(1, 2))  # isolated dst stays exactly zero


@pytest.mark.parametrize("name", GT_CASES)
def test_golden_bf16(name):
    """bf16 storage, fp32 math: compare with the fp32 reference evaluated on the bf16-rounded inputs."""
    z = load_golden(name)
    size = tuple(int(x) for x in z["size"])
    rb = lambda a: t(a).bfloat16().float()
    ref = og.gt_conv_unfused_fwd_bwd(rb(z["q"]), rb(z["k"]), rb(z["v"]), rb(z["e"]), t(z["edge_index"]), rb(z["g"]), size)
    r = run_b2(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), t(z["g"]), size, torch.bfloat16)
    for key in ("out", "dq", "dk", "dv", "de"):
        assert r[key].dtype == torch.bfloat16
        assert rel_err(r[key].float(), ref[key]) < BF16_TOL, (name, key, rel_err(r[key].float(), ref[key]))


def _random_case(seed, ns, nd, E, H, C, zipf=False):
    gen = torch.Generator().manual_seed(seed)
    src = torch.randint(0, ns, (E,), generator=gen)
    if zipf:  # skewed in-degree: a few dst nodes collect most edges
        dst = (torch.rand(E, generator=gen) ** 4 * nd).long().clamp_(max=nd - 1)
    else:
        dst = torch.randint(0, nd, (E,), generator=gen)
    ei = torch.stack([src, dst])
    q = torch.randn(nd, H, C, generator=gen)
    k = torch.randn(ns, H, C, generator=gen)
    v = torch.randn(ns, H, C, generator=gen)
    e = torch.randn(E, H, C, generator=gen)
    g = torch.randn(nd, H, C, generator=gen)
    return q, k, v, e, ei, g


# (H, C) pairs chosen to hit every lanes-per-head specialisation for fp32 (C*4/16) and bf16 (C*2/16), the
# head-sliced launch (H*LPH > 128 threads) and the generic any-C kernels
SHAPES = [(16, 64), (16, 16), (4, 8), (2, 4), (8, 32), (4, 128), (2, 256), (16, 128), (3, 5), (2, 24), (1, 40),
          # 2 KB rows (bulk-copy pipelined kernels): every lanes-per-head value
          (8, 128), (32, 32), (32, 16), (64, 16), (128, 8), (16, 32)]


@pytest.mark.parametrize("H,C", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("E", [2500, 700])  # mean in-degree 16.7 (one dst row per thread group) and 4.7 (row-block kernels)
def test_random_graphs_all_shapes(H, C, dtype, E):
    q, k, v, e, ei, g = _random_case(H * 1000 + C + E, ns=300, nd=150, E=E, H=H, C=C)
    cast = (lambda x: x) if dtype == torch.float32 else (lambda x: x.bfloat16().float())
    ref = og.gt_conv_unfused_fwd_bwd(cast(q), cast(k), cast(v), cast(e), ei, cast(g), (300, 150))
    r = run_b2(q, k, v, e, ei, g, (300, 150), dtype)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key].float(), ref[key]) < tol, (H, C, dtype, key, rel_err(r[key].float(), ref[key]))


@pytest.mark.parametrize("dtype,H,C", [(torch.bfloat16, 16, 64), (torch.float32, 16, 32)])
@pytest.mark.parametrize("nd,E", [(1, 5), (37, 900), (700, 9000), (5000, 30000)])
def test_pipelined_kernel_row_shapes(dtype, H, C, nd, E):
    """2 KB rows take the warp-specialised bulk-copy kernel: skewed in-degrees (rows with no edges, rows longer than the
    32-edge index batch), dst counts below / around / above the number of CTAs, shuffled edge order."""
    gen = torch.Generator().manual_seed(nd + E)
    ns = 400
    dst = (torch.rand(E, generator=gen) ** 3 * nd).long().clamp_(max=nd - 1)
    src = torch.randint(0, ns, (E,), generator=gen)
    ei = torch.stack([src, dst])
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns, E, nd))
    cast = (lambda x: x) if dtype == torch.float32 else (lambda x: x.bfloat16().float())
    ref = og.gt_conv_unfused_fwd_bwd(cast(q), cast(k), cast(v), cast(e), ei, cast(g), (ns, nd))
    r = run_b2(q, k, v, e, ei, g, (ns, nd), dtype)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key].float(), ref[key]) < tol, (key, rel_err(r[key].float(), ref[key]))


def test_skewed_degrees_and_f64_oracle():
    q, k, v, e, ei, g = _random_case(7, ns=500, nd=200, E=20000, H=4, C=16, zipf=True)
    r = run_b2(q, k, v, e, ei, g, None)
    o = og.gt_conv_csr_f64(q, k, v, e, ei, g)
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], torch.from_numpy(o[key])) < FP32_TOL, key


def test_decoder_like_graph_high_out_degree():
    """3 incoming edges per dst from nearby src (decoder shape): src out-degree ~36, many consecutive CSC segments,
    plus a block of src rows without any edge (zero dk/dv rows written by the src pass)."""
    gen = torch.Generator().manual_seed(23)
    ns, nd, H, C = 500, 6000, 4, 16
    base = (torch.arange(nd) * (ns - 100) // nd)
    src = (base.view(-1, 1) + torch.randint(0, 3, (nd, 3), generator=gen)).clamp_(max=ns - 101).view(-1)  # rows >= 400 unused
    dst = torch.arange(nd).repeat_interleave(3)
    ei = torch.stack([src, dst])
    E = ei.shape[1]
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns, E, nd))
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns, nd))
    r = run_b2(q, k, v, e, ei, g, (ns, nd))
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], ref[key]) < FP32_TOL, (key, rel_err(r[key], ref[key]))
    assert float(r["dk"][400:].abs().max()) == 0.0 and float(r["dv"][400:].abs().max()) == 0.0
    rb = run_b2(q, k, v, e, ei, g, (ns, nd), torch.bfloat16)
    cast = lambda x: x.bfloat16().float()
    refb = og.gt_conv_unfused_fwd_bwd(cast(q), cast(k), cast(v), cast(e), ei, cast(g), (ns, nd))
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(rb[key].float(), refb[key]) < BF16_TOL, (key, rel_err(rb[key].float(), refb[key]))


@pytest.mark.parametrize("nd", [1, 3, 4, 5, 9])
def test_dst_row_block_boundaries(nd):
    """Low in-degree graphs: a thread group owns 4 consecutive dst rows; row counts around the block size, rows
    without edges at the start, in the middle and at the end of a block."""
    gen = torch.Generator().manual_seed(100 + nd)
    ns, H, C = 13, 2, 8
    E = 3 * nd
    dst = torch.randint(0, nd, (E,), generator=gen)
    if nd >= 4:
        dst[dst == 0] = nd - 2  # row 0 and the last row stay empty
        dst[dst == nd - 1] = 1
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), dst])
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns, E, nd))
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns, nd))
    r = run_b2(q, k, v, e, ei, g, (ns, nd))
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], ref[key]) < FP32_TOL, (key, rel_err(r[key], ref[key]))


@pytest.mark.parametrize("ns", [1, 7, 8, 9, 17])
def test_src_row_block_boundaries(ns):
    """The src pass owns blocks of 8 consecutive src rows: exercise row counts around the block size."""
    gen = torch.Generator().manual_seed(ns)
    nd, E, H, C = 11, 60, 2, 8
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd, (E,), generator=gen)])
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns, E, nd))
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns, nd))
    r = run_b2(q, k, v, e, ei, g, (ns, nd))
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], ref[key]) < FP32_TOL, (key, rel_err(r[key], ref[key]))


def test_empty_edge_set_and_size_errors():
    import anemoi_models_b200 as b2

    conv = b2.GraphTransformerConv(out_channels=8)
    q = torch.randn(5, 2, 8, device="cuda", requires_grad=True)
    k = torch.randn(7, 2, 8, device="cuda", requires_grad=True)
    e = torch.zeros(0, 2, 8, device="cuda", requires_grad=True)
    ei = torch.zeros(2, 0, dtype=torch.long, device="cuda")
    out = conv(q, k, k, e, ei)
    assert out.shape == (5, 2, 8) and float(out.abs().max()) == 0.0
    out.sum().backward()
    assert float(q.grad.abs().max()) == 0.0 and float(k.grad.abs().max()) == 0.0 and e.grad.shape == (0, 2, 8)
    with pytest.raises(ValueError):
        conv(q, k, k, e, ei, size=(8, 5))
    with pytest.raises(ValueError):
        conv(q, k, k, e, ei.float())
    with pytest.raises(TypeError):
        conv(q, k, k, None, ei)


def test_non_contiguous_views_and_no_grad_inputs():
    """Blocks hand the conv einops/reshape views; some inputs may not require grad (ctx.needs_input_grad)."""
    import anemoi_models_b200 as b2

    q, k, v, e, ei, g = _random_case(11, ns=60, nd=40, E=400, H=4, C=8)
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (60, 40))
    conv = b2.GraphTransformerConv(out_channels=8)
    qc = q.cuda().transpose(0, 1).contiguous().transpose(0, 1).requires_grad_(True)  # non-contiguous view
    kc, vc = k.cuda(), v.cuda().requires_grad_(True)
    ec = e.cuda().requires_grad_(True)
    out = conv(qc, kc, vc, ec, ei.cuda(), (60, 40))
    out.backward(g.cuda())
    assert rel_err(out, ref["out"]) < FP32_TOL and rel_err(qc.grad, ref["dq"]) < FP32_TOL
    assert kc.grad is None and rel_err(vc.grad, ref["dv"]) < FP32_TOL and rel_err(ec.grad, ref["de"]) < FP32_TOL


def test_edge_order_invariance_is_bit_exact():
    """out and the node gradients depend on the edge SET only: a shuffled edge list (perm != identity) gives
    results equal to the dst-sorted list up to summation order inside a segment (stable sort keeps it fixed)."""
    q, k, v, e, ei, g = _random_case(13, ns=200, nd=100, E=3000, H=4, C=16)
    order = torch.sort(ei[1], stable=True).indices
    a = run_b2(q, k, v, e, ei, g, (200, 100))
    b = run_b2(q, k, v, e[order], ei[:, order], g, (200, 100))
    assert torch.equal(a["out"], b["out"]) and torch.equal(a["dq"], b["dq"])
    assert torch.equal(a["de"][order], b["de"])
    assert torch.equal(a["dk"], b["dk"]) and torch.equal(a["dv"], b["dv"])


def test_checkpoint_recompute_and_determinism():
    import anemoi_models_b200 as b2
    from torch.utils.checkpoint import checkpoint

    q, k, v, e, ei, g = _random_case(17, ns=80, nd=50, E=600, H=2, C=16)
    conv = b2.GraphTransformerConv(out_channels=16)
    ins = [x.cuda().requires_grad_(True) for x in (q, k, v, e)]
    out = checkpoint(lambda *a: conv(*a, ei.cuda(), (80, 50)), *ins, use_reentrant=False)
    out.backward(g.cuda())
    plain = run_b2(q, k, v, e, ei, g, (80, 50))
    for x, key in zip(ins, ("dq", "dk", "dv", "de")):
        assert torch.equal(x.grad, plain[key])


def test_headline_shape_properties_bf16():
    """Full-size n320->o96-like shapes (E ~ 750k, D = 1024, bf16): size-independent properties.
    (1) softmax weights sum to one: with v = c (constant rows) and e = 0 the output of every non-isolated dst is c;
    (2) linearity in (v, e-value path): out(v1+v2) = out(v1)+out(v2) when the logits are unchanged (k fixed, e = 0);
    (3) conservation: column sums of dv over src equal column sums of g over dst (weights of a segment sum to one)."""
    import anemoi_models_b200 as b2

    torch.manual_seed(0)
    Ns, Nd, H, C = 542080, 40320, 16, 64
    deg = torch.randint(16, 23, (Nd,))
    dst = torch.repeat_interleave(torch.arange(Nd), deg)
    E = dst.numel()
    src = (dst * (Ns // Nd) + torch.randint(-40, 40, (E,))).clamp_(0, Ns - 1)
    ei = torch.stack([src, dst]).cuda()
    conv = b2.GraphTransformerConv(out_channels=C)
    q = torch.randn(Nd, H, C, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(Ns, H, C, device="cuda", dtype=torch.bfloat16)
    e0 = torch.zeros(E, H, C, device="cuda", dtype=torch.bfloat16)
    vconst = torch.full((Ns, H, C), 0.75, device="cuda", dtype=torch.bfloat16)
    out = conv(q, k, vconst, e0, ei, (Ns, Nd))
    assert float((out.float() - 0.75).abs().max()) < 1e-2
    v1 = torch.randn(Ns, H, C, device="cuda", dtype=torch.bfloat16)
    v2 = torch.randn(Ns, H, C, device="cuda", dtype=torch.bfloat16)
    o1, o2 = conv(q, k, v1, e0, ei, (Ns, Nd)).float(), conv(q, k, v2, e0, ei, (Ns, Nd)).float()
    o12 = conv(q, k, (v1.float() + v2.float()).bfloat16(), e0, ei, (Ns, Nd)).float()
    assert float((o12 - (o1 + o2)).abs().max()) < 6e-2
    v1.requires_grad_(True)
    g = torch.randn(Nd, H, C, device="cuda", dtype=torch.bfloat16)
    conv(q, k, v1, e0, ei, (Ns, Nd)).backward(g)
    # every column: sum_j dv[j,h,c] = sum_i g[i,h,c] * (sum_t a_t) = sum_i g[i,h,c]   (bf16 rounding of dv: ~0.5 abs, 6 sigma = 3)
    col_dv = v1.grad.float().sum(0)
    col_g = g.float().sum(0)
    assert float((col_dv - col_g).abs().max()) < 3.0


@pytest.mark.parametrize("which", ["enc", "proc", "dec"])
def test_config1_graphs_full_size_fp32(which):
    """BASELINE configs[0] (the reference's own CPU-runnable case): o96 (40,320) data grid, o48 (10,944) hidden grid,
    D = 256, 16 heads, fp32 -- the three conv shapes of that model at FULL size against the CPU oracle
    (encoder cut-off 0.6: E = 51,608; processor 8-NN: 87,552; decoder 3-NN: 120,960)."""
    from anemoi_models_b200 import synthetic as S

    data, _ = S.octahedral_grid(96)
    hidden, _ = S.octahedral_grid(48)
    if which == "enc":
        ei_np = S.cutoff_edges(data, hidden, 0.6 * S.max_nn_distance(hidden))
        ns, nd, expect = len(data), len(hidden), 51608
    elif which == "proc":
        ei_np = S.knn_edges(hidden, hidden, 8, exclude_self=True)
        ns, nd, expect = len(hidden), len(hidden), 87552
    else:
        ei_np = S.knn_edges(hidden, data, 3)
        ns, nd, expect = len(hidden), len(data), 120960
    assert (ns, nd) in ((40320, 10944), (10944, 10944), (10944, 40320))
    E = ei_np.shape[1]
    assert abs(E - expect) <= 0.01 * expect, (which, E)  # SURVEY 8d edge counts (the cut-off set depends on the grid recipe)
    gen = torch.Generator().manual_seed(0)
    H, C = 16, 16
    ei = torch.from_numpy(ei_np)
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen) for n in (nd, ns, ns, E, nd))
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns, nd))
    r = run_b2(q, k, v, e, ei, g, (ns, nd))
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], ref[key]) < FP32_TOL, (which, key, rel_err(r[key], ref[key]))


def test_conv_step_is_cuda_graph_capturable():
    """The C-ABI calls only enqueue work on the caller's stream (no allocation, no synchronisation), so a whole
    forward+backward of the conv can be captured in a CUDA graph and replayed -- what a launch-bound small model
    (BASELINE configs[0] shapes: ~0.2 ms per conv) wants."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200.graph import get_csr

    q, k, v, e, ei, g = _random_case(29, ns=400, nd=250, E=3000, H=16, C=16)
    ref = og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (400, 250))
    conv = b2.GraphTransformerConv(out_channels=16)
    ei_c = ei.cuda()
    plan = get_csr(ei_c, 400, 250)  # the plan is built (and cached) outside the capture: that step synchronises once
    ins = [x.cuda().requires_grad_(True) for x in (q, k, v, e)]
    gc = g.cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up outside the capture
        for _ in range(2):
            for x in ins:
                x.grad = None
            conv(*ins, ei_c, (400, 250), plan=plan).backward(gc)
    torch.cuda.current_stream().wait_stream(side)
    for x in ins:
        x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = conv(*ins, ei_c, (400, 250), plan=plan)
        out.backward(gc)
    for x in ins:  # new values in the static input buffers, then replay
        x.data.mul_(-1.0)
    graph.replay()
    torch.cuda.synchronize()
    ref2 = og.gt_conv_unfused_fwd_bwd(-q, -k, -v, -e, ei, g, (400, 250))
    assert rel_err(out, ref2["out"]) < FP32_TOL
    for x, key in zip(ins, ("dq", "dk", "dv", "de")):
        assert rel_err(x.grad, ref2[key]) < FP32_TOL, key
    assert rel_err(ref["out"], ref["out"]) == 0.0


def test_rare_configurations_run_as_torch_compositions_on_the_gpu(monkeypatch):
    """attention dropout > 0 in training mode and GraphConv activations outside SiLU / GELU / ReLU / Identity used to
    raise; they now run the reference's op sequence as CUDA torch ops.  fp16 inputs come back as fp16; a contiguous view at a
    storage offset that is not 16-byte aligned is handled."""
    import anemoi_models_b200 as b2

    torch.manual_seed(0)
    ns, nd, E, H, C = 50, 30, 400, 4, 8
    ei = torch.stack([torch.randint(0, ns, (E,)), torch.randint(0, nd, (E,))]).cuda()
    q, k, v, e = (torch.randn(n, H, C, device="cuda") for n in (nd, ns, ns, E))
    fused = b2.GraphTransformerConv(C)(q, k, v, e, ei, (ns, nd))
    conv = b2.GraphTransformerConv(C, dropout=0.3).train()
    monkeypatch.setattr(torch.nn.functional, "dropout", lambda x, p, training: x)  # identity: the composition must equal the fused conv
    assert rel_err(conv(q, k, v, e, ei, (ns, nd)), fused) < 1e-5
    monkeypatch.undo()
    out = conv(q, k, v, e, ei, (ns, nd))
    assert out.shape == fused.shape and bool(torch.isfinite(out).all()) and not torch.allclose(out, fused)
    assert rel_err(conv.eval()(q, k, v, e, ei, (ns, nd)), fused) == 0.0  # eval mode: the fused kernels
    # fp16 in -> fp16 out (computed in fp32)
    out16 = b2.GraphTransformerConv(C)(q.half(), k.half(), v.half(), e.half(), ei, (ns, nd))
    assert out16.dtype == torch.float16 and rel_err(out16.float(), fused) < 2e-3
    # misaligned contiguous view (storage offset 1 element = 4 bytes)
    buf = torch.randn(nd * H * C + 1, device="cuda")
    q_off = buf[1:].view(nd, H, C)
    q_off.copy_(q)
    assert q_off.data_ptr() % 16 != 0 and q_off.is_contiguous()
    assert rel_err(b2.GraphTransformerConv(C)(q_off, k, v, e, ei, (ns, nd)), fused) == 0.0
    # GraphConv with an activation the fused kernels do not implement
    from oracle import gtconv as og

    D = 16
    gc = b2.GraphConv(D, D, activation="Tanh").cuda()
    x, ea = torch.randn(ns, D), torch.randn(E, D)
    ei2 = torch.stack([ei[0].cpu(), ei[1].cpu() % ns])
    p = {kk: vv.detach().cpu() for kk, vv in gc.state_dict().items()}
    o_ref, en_ref = og.graph_conv_unfused(x, ea, ei2, p, "edge_mlp.", 0, "Tanh")
    o, en = gc(x.cuda(), ea.cuda(), ei2.cuda())
    assert rel_err(o, o_ref) < 2e-5 and rel_err(en, en_ref) < 2e-5


def test_high_out_degree_src_pass_interleaved_rows_bf16():
    """Decoder-like graph (3 src per dst, mean out-degree 40 -> the pipelined src pass deals the src rows to the CTAs interleaved,
    csrc/gtconv_tma.cu) with 2 KB rows: edge-less src rows, one src row with > 64 outgoing edges (index batches reloaded on
    demand) and more rows than one pointer batch per CTA would cover, against the fp32 op sequence on bf16-rounded inputs."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200 import _lib

    gen = torch.Generator().manual_seed(5)
    ns, nd, H, C = 151, 2011, 16, 64
    base = (torch.arange(nd) * (ns - 9) // nd)
    src = torch.stack([base, base + 1, base + 2 + torch.randint(0, 6, (nd,), generator=gen)], 1).clamp_(max=ns - 1).view(-1)
    dst = torch.arange(nd).repeat_interleave(3)
    keep = (src != 40) & (src != 41)                      # two src rows without any edge
    src, dst = src[keep], dst[keep]
    extra_dst = torch.randperm(nd, generator=gen)[:90]    # src row 7 gets 90 more edges (> 64: two index batches are not enough)
    src, dst = torch.cat([src, torch.full((90,), 7)]), torch.cat([dst, extra_dst])
    perm = torch.randperm(src.numel(), generator=gen)
    ei = torch.stack([src, dst])[:, perm].contiguous()
    E = ei.shape[1]
    assert E >= 20 * ns
    name = _lib.lib().ab2_gtconv_variant(2, 1, ns, nd, E, H, C).decode()
    assert "src_tma" in name, name
    q, k, v, e, g = (torch.randn(n, H, C, generator=gen).bfloat16() for n in (nd, ns, ns, E, nd))
    ref_in = [t_.float().requires_grad_(True) for t_ in (q, k, v, e)]
    ref = og.gt_conv_unfused(*ref_in, ei, (ns, nd))
    ref.backward(g.float())
    ins = [t_.cuda().requires_grad_(True) for t_ in (q, k, v, e)]
    out = b2.GraphTransformerConv(C)(*ins, ei.cuda(), (ns, nd))
    out.backward(g.cuda())
    assert rel_err(out.float(), ref) < 2e-2
    for nm, a_, b_ in zip("qkve", ins, ref_in):
        assert rel_err(a_.grad.float(), b_.grad) < 2e-2, nm
    assert float(ins[1].grad[40:42].abs().max()) == 0.0 and float(ins[2].grad[40:42].abs().max()) == 0.0
