"""Like-for-like comparator: the reference's unfused op sequence (oracle port of conv.py + PyG, plain torch
ops) executed ON THE GPU against the fused CUDA path, same inputs, headline-like shapes.  Checks parity at full size and
prints both timings (informational; the numbers quoted in profiles/ come from bench.py)."""
import time

import pytest
import torch

from conftest import rel_err
from oracle import gtconv as og

pytestmark = pytest.mark.gpu


def _time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_matches_unfused_on_gpu_at_scale(dtype, capsys):
    import anemoi_models_b200 as b2

    torch.manual_seed(0)
    # 1/4 of the headline graph in every dimension: the unfused sequence keeps ~10 [E, D] temporaries alive
    Ns, Nd, H, C = 135520, 10080, 16, 64
    deg = torch.randint(16, 23, (Nd,))
    dst = torch.repeat_interleave(torch.arange(Nd), deg)
    E = dst.numel()
    src = (dst * (Ns // Nd) + torch.randint(-40, 40, (E,))).clamp_(0, Ns - 1)
    ei = torch.stack([src, dst]).cuda()
    q, k, v, e, g = (torch.randn(n, H, C, device="cuda").to(dtype) for n in (Nd, Ns, Ns, E, Nd))
    conv = b2.GraphTransformerConv(out_channels=C)

    def fused():
        ins = [x.detach().requires_grad_(True) for x in (q, k, v, e)]
        out = conv(*ins, ei, (Ns, Nd))
        out.backward(g)
        return out, ins

    def unfused():
        ins = [x.detach().float().requires_grad_(True) for x in (q, k, v, e)]  # fp32 math on the same (rounded) inputs
        out = og.gt_conv_unfused(*ins, ei, (Ns, Nd))
        out.backward(g.float())
        return out, ins

    out_f, ins_f = fused()
    out_u, ins_u = unfused()
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel_err(out_f.float(), out_u) < tol
    for a, b_, name in zip(ins_f, ins_u, "qkve"):
        assert rel_err(a.grad.float(), b_.grad) < tol, name
    t_f, t_u = _time(fused), _time(unfused, 3)
    with capsys.disabled():
        print(f"\n[compare {dtype}] E={E}: fused {t_f:.3f} ms ({E / t_f / 1e3:.1f} M edges/s)  |  reference op sequence on the same GPU "
              f"(fp32 torch ops) {t_u:.3f} ms ({E / t_u / 1e3:.1f} M edges/s)  |  speed-up x{t_u / t_f:.1f}")
    assert t_f < t_u
