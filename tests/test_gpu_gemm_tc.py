"""tcgen05 / TMEM GEMM (csrc/gemm_tc.cu), LayerNorm and column-sum kernels through the C ABI, against fp32 torch on the same
bf16-rounded inputs (the "plain PyTorch fp32 reference of the same op"), and the GT blocks running on them against the same
blocks on nn.Linear / nn.LayerNorm.  Tolerances: bf16 results <= 2e-2 of max|ref| and <= 1e-2 relative L2; fp32 outputs
(split-K wgrad partial sums) <= 1e-4."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mk(M, N, K, seed=0):
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=gen).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=gen).to(DEV).bfloat16()
    return A, B, A.float() @ B.float().t()


@pytest.mark.parametrize("cg", ["2", "1"])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1000, 264, 200), (300, 128, 256), (77, 64, 72), (4096, 2048, 1024), (1, 8, 8)])
def test_gemm_layouts_and_tails(M, N, K, cg, monkeypatch):
    """K-major / MN-major operand combinations (forward, dgrad, wgrad layouts), ragged M / N / K, CTA pairs and single CTAs."""
    from anemoi_models_b200 import gemm as G

    monkeypatch.setenv("AB2_GEMM_CG", cg)
    A, B, ref = _mk(M, N, K)
    for a_mn, b_mn in ((False, False), (False, True), (True, True), (True, False)):
        if (a_mn and M % 8) or (b_mn and N % 8):
            continue  # an MN-major operand needs a 16-byte aligned row stride
        a_arg = A.t().contiguous() if a_mn else A
        b_arg = B.t().contiguous() if b_mn else B
        out = G.gemm(a_arg, b_arg, M, N, K, a_mn=a_mn, b_mn=b_mn)
        assert rel_err(out, ref) < 1e-2 and rel_l2(out, ref) < 5e-3, (a_mn, b_mn)
        out32 = G.gemm(a_arg, b_arg, M, N, K, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
        assert rel_err(out32, ref) < 1e-5, (a_mn, b_mn)


@pytest.mark.parametrize("splits", [2, 5, 8])
def test_split_k_wgrad_is_deterministic(splits):
    from anemoi_models_b200 import gemm as G

    M, N, K = 264, 200, 4000  # dW [264, 200] = g^T x over 4000 rows
    A, B, ref = _mk(M, N, K, seed=3)
    a_arg, b_arg = A.t().contiguous(), B.t().contiguous()
    o1 = G.gemm(a_arg, b_arg, M, N, K, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
    o2 = G.gemm(a_arg, b_arg, M, N, K, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
    assert torch.equal(o1, o2)
    assert rel_err(o1, ref) < 1e-4
    ob = G.gemm(a_arg, b_arg, M, N, K, a_mn=True, b_mn=True, out_dtype=torch.bfloat16, splits=splits)
    assert rel_err(ob, ref) < 1e-2


def test_gemm_epilogues():
    from anemoi_models_b200 import gemm as G

    M, N, K = 1000, 512, 256
    A, B, acc = _mk(M, N, K, seed=5)
    gen = torch.Generator().manual_seed(6)
    bias = torch.randn(N, generator=gen).to(DEV)
    res16 = torch.randn(M, N, generator=gen).to(DEV).bfloat16()
    res32 = torch.randn(M, N, generator=gen).to(DEV)
    # bias + residual, bf16 and fp32 residual (output takes the residual's dtype in the blocks)
    assert rel_err(G.gemm(A, B, M, N, K, bias=bias, residual=res16), acc + bias + res16.float()) < 1e-2
    out = G.gemm(A, B, M, N, K, bias=bias, residual=res32, out_dtype=torch.float32)
    assert rel_err(out, acc + bias + res32) < 1e-5
    # activation with the pre-activation side output (the activation is applied to the bf16-rounded pre-activation)
    for name, fn in (("GELU", F.gelu), ("SiLU", F.silu), ("ReLU", F.relu)):
        pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        h = G.gemm(A, B, M, N, K, bias=bias, act=G.ACT_CODES[name], pre_out=pre)
        assert rel_err(pre, acc + bias) < 1e-2
        assert rel_err(h, fn(pre.float())) < 5e-3, name
        # dgrad through the activation: out = (A B^T) * act'(pre)
        x = pre.float().requires_grad_(True)
        fn(x).sum().backward()
        dg = G.gemm(A, B, M, N, K, dact_pre=pre, act=G.ACT_CODES[name])
        assert rel_err(dg, acc * x.grad) < 1e-2, name
    # LayerNorm folded into the GEMM: row_scale * acc + row_shift * col_vec
    rs, rt, cv = torch.rand(M, generator=gen).to(DEV) + 0.5, torch.randn(M, generator=gen).to(DEV), torch.randn(N, generator=gen).to(DEV)
    out = G.gemm(A, B, M, N, K, row_scale=rs, row_shift=rt, col_vec=cv, bias=bias)
    assert rel_err(out, rs[:, None] * acc + rt[:, None] * cv[None, :] + bias) < 1e-2
    # column segments: q | self land in separate tensors
    outs = [torch.empty(M, N // 2, device=DEV, dtype=torch.bfloat16) for _ in range(2)]
    G.gemm(A, B, M, N, K, bias=bias, out=outs, seg_cols=N // 2)
    assert rel_err(torch.cat(outs, 1), acc + bias) < 1e-2


@pytest.mark.parametrize("nseg", [2, 3, 4])
def test_segmented_a_operand_matches_the_concatenated_one(nseg):
    """A given as 2-4 separate tensors (the q | k | v | self gradients where autograd left them): K-major segments along the
    contraction (fused dgrad) and MN-major segments along the output rows (fused split-K wgrad) give the same result, bit for
    bit, as the same GEMM on the concatenated operand."""
    from anemoi_models_b200 import gemm as G

    torch.manual_seed(nseg)
    M, Nw, Kin = 1000, 256, 200  # rows, width of one layer, input width
    dys = [torch.randn(M, Nw, device=DEV).bfloat16() for _ in range(nseg)]
    w = torch.randn(nseg * Nw, Kin, device=DEV).bfloat16()
    x = torch.randn(M, Kin, device=DEV).bfloat16()
    cat = torch.cat(dys, dim=1)
    dx_seg = G.gemm(dys[0], w, M, Kin, nseg * Nw, b_mn=True, a_segs=dys[1:], a_seg_len=Nw)
    dx_cat = G.gemm(cat, w, M, Kin, nseg * Nw, b_mn=True)
    assert torch.equal(dx_seg, dx_cat)
    assert rel_err(dx_seg, cat.float() @ w.float()) < 1e-2
    for splits in (1, 3):
        dw_seg = G.gemm(dys[0], x, nseg * Nw, Kin, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits, a_segs=dys[1:],
                        a_seg_len=Nw)
        dw_cat = G.gemm(cat, x, nseg * Nw, Kin, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
        assert torch.equal(dw_seg, dw_cat)
        assert rel_err(dw_seg, cat.float().t() @ x.float()) < 1e-4
    with pytest.raises(ValueError):  # segment width must be a multiple of 64 (K-major)
        G.gemm(dys[0][:, :40].contiguous(), w, M, Kin, 80, b_mn=True, a_segs=[dys[1][:, :40].contiguous()], a_seg_len=40)
    with pytest.raises(ValueError):  # MN-major segments must be whole 256-row tiles
        G.gemm(dys[0][:, :128].contiguous(), x, 256, Kin, M, a_mn=True, b_mn=True, a_segs=[dys[1][:, :128].contiguous()], a_seg_len=128)


def test_linear_multi_and_layernorm_fork_autograd():
    """The GT block's attention input: (LayerNorm(x), skip) fork and q | k | v | self as one GEMM forward / one dgrad / one
    wgrad, against the same layers run one by one through torch fp32 on bf16-rounded parameters."""
    from anemoi_models_b200 import gemm as G

    torch.manual_seed(0)
    M, D = 777, 256
    ln = torch.nn.LayerNorm(D).to(DEV)
    lins = [torch.nn.Linear(D, D).to(DEV) for _ in range(4)]
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
    x = torch.randn(M, D, device=DEV)
    gs = [torch.randn(M, D, device=DEV).bfloat16() for _ in range(4)] + [torch.randn(M, D, device=DEV)]
    mods = [ln] + lins
    for m in mods:
        m.zero_grad()
    x1 = x.clone().requires_grad_(True)
    xn, skip = G.layer_norm_fork(x1, ln)
    ys = G.linear_multi(xn, lins)
    torch.autograd.backward(list(ys) + [skip], gs)
    got = list(ys) + [x1.grad] + [p.grad.clone() for m in mods for p in m.parameters()]
    for m in mods:
        m.zero_grad()
    xr = x.clone().requires_grad_(True)
    xnr = ln(xr)
    yr = [l(xnr) for l in lins]  # plain fp32 modules: the bf16 rounding of LN(x) and of the weights is inside the tolerance
    torch.autograd.backward(yr + [xr], [g.float() for g in gs])
    ref = yr + [xr.grad] + [p.grad for m in mods for p in m.parameters()]
    for i, (a, b) in enumerate(zip(got, ref)):
        assert rel_err(a, b) < 2e-2 and rel_l2(a, b) < 2e-2, (i, rel_err(a, b), rel_l2(a, b))
    # a second call gives bit-identical gradients (fixed-order split-K partial sums, no atomics)
    for m in mods:
        m.zero_grad()
    x2 = x.clone().requires_grad_(True)
    xn2, skip2 = G.layer_norm_fork(x2, ln)
    torch.autograd.backward(list(G.linear_multi(xn2, lins)) + [skip2], gs)
    assert torch.equal(x2.grad, got[4]) and torch.equal(lins[2].weight.grad, got[5 + 2 + 2 * 2])


def test_gemm_argument_errors():
    from anemoi_models_b200 import gemm as G

    A, B, _ = _mk(64, 64, 64)
    with pytest.raises(ValueError):
        G.gemm(A, B[:, :60].contiguous(), 64, 60, 64)  # N % 8
    with pytest.raises(TypeError):
        G.gemm(A.float(), B, 64, 64, 64)
    with pytest.raises(RuntimeError):
        G.gemm(A.cpu(), B.cpu(), 64, 64, 64)


@pytest.mark.parametrize("xd,yd", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)])
@pytest.mark.parametrize("M,D", [(1000, 1024), (77, 256), (5, 8), (300, 520), (123, 2048), (2500, 1024)])
def test_layernorm_forward_backward(M, D, xd, yd):
    from anemoi_models_b200 import gemm as G

    gen = torch.Generator().manual_seed(M + D)
    x = (torch.randn(M, D, generator=gen) * 2 + 0.5).to(DEV).to(xd)
    ln = torch.nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        ln.weight.copy_(torch.rand(D, generator=gen) + 0.5)
        ln.bias.copy_(torch.randn(D, generator=gen))
    g = torch.randn(M, D, generator=gen).to(DEV)
    xr = x.detach().float().clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), ln.weight, ln.bias, ln.eps)
    ref.backward(g)
    ref_dw, ref_db = ln.weight.grad.clone(), ln.bias.grad.clone()
    ln.zero_grad()
    x1 = x.detach().clone().requires_grad_(True)
    y = G.layer_norm(x1, ln, out_dtype=yd)
    assert y.dtype == yd
    y.backward(g.to(yd))
    tol = 1e-5 if (xd == torch.float32 and yd == torch.float32) else 2e-2
    assert rel_err(y, ref) < tol
    assert x1.grad.dtype == xd and rel_err(x1.grad, xr.grad) < tol and rel_l2(x1.grad, xr.grad) < tol
    assert rel_err(ln.weight.grad, ref_dw) < tol and rel_err(ln.bias.grad, ref_db) < tol
    # deterministic
    ln.zero_grad()
    x2 = x.detach().clone().requires_grad_(True)
    G.layer_norm(x2, ln, out_dtype=yd).backward(g.to(yd))
    assert torch.equal(x2.grad, x1.grad)


def test_layernorm_backward_warp_per_row_variant_in_a_subprocess():
    """AB2_LN_BWD=rows keeps the first backward kernel (one warp per row, accumulators for all of D in registers) for A/B
    runs; the switch is read once per process, so the kept kernel is checked in a child process."""
    import os
    import subprocess
    import sys

    code = (
        "import torch, torch.nn.functional as F\n"
        "from anemoi_models_b200 import gemm as G\n"
        "torch.manual_seed(0)\n"
        "x = torch.randn(777, 1024, device='cuda'); g = torch.randn(777, 1024, device='cuda')\n"
        "ln = torch.nn.LayerNorm(1024).to('cuda')\n"
        "xr = x.clone().requires_grad_(True); F.layer_norm(xr, (1024,), ln.weight, ln.bias, ln.eps).backward(g)\n"
        "rw = ln.weight.grad.clone(); ln.zero_grad()\n"
        "x1 = x.clone().requires_grad_(True); G.layer_norm(x1, ln, out_dtype=torch.float32).backward(g)\n"
        "e1 = float((x1.grad - xr.grad).abs().max() / xr.grad.abs().max()); e2 = float((ln.weight.grad - rw).abs().max() / rw.abs().max())\n"
        "assert e1 < 1e-5 and e2 < 1e-5, (e1, e2)\n"
        "print('ok')\n")
    env = dict(os.environ, AB2_LN_BWD="rows")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=root)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-800:]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,k", [(1000, 11), (7, 3), (513, 16), (100, 24)])
def test_pad_cast_rows_and_its_backward(rows, k, dt):
    """Raw edge features -> zero-padded bf16 rows for the lin_edge GEMM (one kernel each way) against F.pad + cast, exact."""
    from anemoi_models_b200 import gemm as G

    torch.manual_seed(rows + k)
    x = torch.randn(rows, k, device=DEV).to(dt).requires_grad_(True)
    y = G.pad_cast(x, 16)
    kp = ((k + 15) // 16) * 16
    assert y.shape == (rows, kp) and y.dtype == torch.bfloat16
    assert torch.equal(y, F.pad(x.detach(), (0, kp - k)).bfloat16())
    g = torch.randn(rows, kp, device=DEV).bfloat16()
    y.backward(g)
    assert x.grad.dtype == dt and torch.equal(x.grad, g[:, :k].to(dt))
    xs = torch.randn(rows, k + 5, device=DEV).to(dt)[:, :k]  # a row stride that is not k
    assert torch.equal(G.pad_cast(xs, 16), F.pad(xs, (0, kp - k)).bfloat16())


def test_colsum():
    from anemoi_models_b200 import gemm as G

    gen = torch.Generator().manual_seed(9)
    for M, N, dt in ((5000, 2048, torch.bfloat16), (3, 8, torch.float32), (1234, 520, torch.bfloat16), (40321, 1024, torch.bfloat16),
                     (9000, 4096, torch.bfloat16), (777, 256, torch.float32)):
        a = torch.randn(M, N, generator=gen).to(DEV).to(dt)
        assert rel_err(G.colsum(a), a.float().sum(0)) < 1e-5
        assert torch.equal(G.colsum(a), G.colsum(a))


def test_linear_chain_autograd():
    """linear -> (activation in the epilogue) -> act_linear + residual, forward and backward, against torch fp32 on the same
    bf16-rounded parameters."""
    from anemoi_models_b200 import gemm as G

    torch.manual_seed(0)
    M, D, Hd = 777, 256, 512
    l1, l2 = torch.nn.Linear(D, Hd).to(DEV), torch.nn.Linear(Hd, D).to(DEV)
    x = torch.randn(M, D, device=DEV).bfloat16()
    g = torch.randn(M, D, device=DEV)
    for name, fn in (("GELU", F.gelu), ("SiLU", F.silu)):
        act = G.ACT_CODES[name]
        for l in (l1, l2):
            l.zero_grad()
        x1 = x.clone().requires_grad_(True)
        pre, h = G.linear(x1, l1, act_out=act)
        y = G.act_linear(pre, h, l2, act, residual=x1)
        y.backward(g.bfloat16())
        got = {"y": y, "dx": x1.grad, "dw1": l1.weight.grad.clone(), "db1": l1.bias.grad.clone(), "dw2": l2.weight.grad.clone(),
               "db2": l2.bias.grad.clone()}
        for l in (l1, l2):
            l.zero_grad()
        xr = x.float().requires_grad_(True)
        w1, w2 = l1.weight.bfloat16().float(), l2.weight.bfloat16().float()
        yr = F.linear(fn(F.linear(xr, w1, l1.bias)), w2, l2.bias) + xr
        yr.backward(g.bfloat16().float())
        # parameter grads of the rounded weights = grads of the parameters (the rounding has derivative one)
        gw1, gb1, gw2, gb2 = torch.autograd.grad(F.linear(fn(F.linear(x.float(), l1.weight, l1.bias)), l2.weight, l2.bias),
                                                 (l1.weight, l1.bias, l2.weight, l2.bias), g.bfloat16().float())
        ref = {"y": yr, "dx": xr.grad, "dw1": gw1, "db1": gb1, "dw2": gw2, "db2": gb2}
        for k in ref:
            assert rel_err(got[k], ref[k]) < 2e-2 and rel_l2(got[k], ref[k]) < 2e-2, (name, k, rel_err(got[k], ref[k]), rel_l2(got[k], ref[k]))


@pytest.mark.parametrize("kind", ["mapper", "processor"])
@pytest.mark.parametrize("mode", ["autocast", "bf16"])
def test_gt_blocks_on_tensor_core_kernels_match_the_library_path(kind, mode):
    """The GT blocks with LayerNorm / Linear on the tcgen05 kernels vs the same blocks on nn.LayerNorm / nn.Linear (cuBLASLt), same
    weights, bf16 autocast (fp32 parameters and residual stream) and a pure bf16 module: outputs, input gradients and every
    parameter gradient.  Both sides compute in bf16, so they are compared with each other at the bf16 tolerance and, for the
    outputs, with the fp32 block on the same inputs."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200 import gemm as G

    torch.manual_seed(1)
    ns, nd, D, H, ed, hid = (700, 300, 256, 8, 11, 512) if kind == "mapper" else (500, 500, 256, 8, 11, 512)
    gen = torch.Generator().manual_seed(2)
    deg = 7
    ei = torch.stack([torch.randint(0, ns, (nd * deg,), generator=gen), torch.arange(nd).repeat_interleave(deg)]).to(DEV)
    E = ei.shape[1]
    cls = b2.GraphTransformerMapperBlock if kind == "mapper" else b2.GraphTransformerProcessorBlock
    blk = cls(D, hid, D, edge_dim=ed, num_heads=H).to(DEV)
    xs, xd = torch.randn(ns, D, generator=gen).to(DEV), torch.randn(nd, D, generator=gen).to(DEV)
    ea, gd = torch.rand(E, ed, generator=gen).to(DEV), torch.randn(nd, D, generator=gen).to(DEV)
    shapes = ([[ns, D]], [[nd, D]], [[E, ed]])

    def run(tc_on, module, dtype):
        G._TC_OFF = not tc_on
        try:
            module.zero_grad()
            a, b, e = (t.to(dtype).clone().requires_grad_(True) for t in (xs, xd, ea))
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "autocast" else torch.autocast("cuda", enabled=False)
            with ctx:
                if kind == "mapper":
                    (_, out), _ = module((a, b), e, ei, shapes, 1, size=(ns, nd))
                else:
                    out, _ = module(a, e, ei, shapes, 1, size=(ns, nd))
            out.backward(gd.to(out.dtype))
            grads = {"d_src": a.grad, "d_edge": e.grad}
            if kind == "mapper":
                grads["d_dst"] = b.grad
            grads.update({n: p.grad.clone() for n, p in module.named_parameters()})
            return out.detach(), grads
        finally:
            G._TC_OFF = False

    if mode == "autocast":
        module, dtype = blk, torch.float32
    else:
        import copy

        module, dtype = copy.deepcopy(blk).bfloat16(), torch.bfloat16
    out_tc, g_tc = run(True, module, dtype)
    out_lib, g_lib = run(False, module, dtype)
    G._TC_OFF = True  # fp32 reference block (no autocast): the ground truth both bf16 paths approximate
    try:
        blk.zero_grad()
        a, b, e = (t.clone().requires_grad_(True) for t in (xs, xd, ea))
        if kind == "mapper":
            (_, out_ref), _ = blk((a, b), e, ei, shapes, 1, size=(ns, nd))
        else:
            out_ref, _ = blk(a, e, ei, shapes, 1, size=(ns, nd))
    finally:
        G._TC_OFF = False
    assert out_tc.dtype == out_lib.dtype
    assert rel_err(out_tc, out_ref) < 2e-2 and rel_l2(out_tc, out_ref) < 2e-2
    assert rel_l2(out_tc, out_ref) < 1.5 * rel_l2(out_lib, out_ref) + 1e-3  # no worse than the library bf16 path
    bad = []
    for k in g_lib:
        assert (g_tc[k] is None) == (g_lib[k] is None), k
        if g_lib[k] is None:
            continue
        e1, e2 = rel_err(g_tc[k], g_lib[k]), rel_l2(g_tc[k], g_lib[k])
        # d lin_key.bias is zero in exact arithmetic (a constant added to every key of a dst shifts all its logits alike and the
        # softmax is invariant to it): both sides hold rounding noise there, only the max-norm bound is meaningful
        if e1 >= 3e-2 or (e2 >= 3e-2 and k != "lin_key.bias"):
            bad.append((k, e1, e2))
    assert not bad, bad
