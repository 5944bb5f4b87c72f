"""The conv with its src rows in two buffers (own shard + halo buffer: what one rank of a dst-row-sharded graph runs) on ONE GPU:
must equal the same conv on the concatenated buffer, outputs and every gradient.  Covers the SPLIT instantiations of every kernel
family (bulk-copy pipelined, LDG, warp-cooperative / per-thread src pass, generic) without needing a second GPU; the multi-GPU
transport itself is covered by tests/test_gpu_multi.py."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,C,dtype,tol", [
    (16, 64, torch.bfloat16, 1e-2),  # 2 KB rows: pipelined forward / dst pass; src pass by out-degree
    (16, 32, torch.float32, 1e-6),   # 2 KB rows, fp32
    (16, 16, torch.float32, 1e-6),   # 1 KB rows: LDG kernels, warp-cooperative src pass
    (4, 8, torch.float32, 1e-6),     # 8 threads per row (several rows per warp): per-thread src pass
    (4, 5, torch.float32, 1e-6),     # generic kernels
])
@pytest.mark.parametrize("deg", [2, 9])  # low / high out-degree: LDG vs pipelined src pass for 2 KB rows
def test_split_src_buffers_equal_one_buffer(H, C, dtype, tol, deg):
    from anemoi_models_b200 import ops
    from anemoi_models_b200.graph import GraphCSR

    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(H * 100 + C + deg)
    nd = 700
    ns, E = (nd * 9 // deg, nd * 9)  # E / ns = deg
    n_own = ns * 3 // 5
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd - 2, (E,), generator=gen)]).to(dev)
    q, g = (torch.randn(nd, H, C, generator=gen).to(dev, dtype) for _ in range(2))
    k, v = (torch.randn(ns, H, C, generator=gen).to(dev, dtype) for _ in range(2))
    e = torch.randn(E, H, C, generator=gen).to(dev, dtype)
    plan = GraphCSR(ei, ns, nd)

    a = [t.clone().requires_grad_(True) for t in (q, k, v, e)]
    ref = ops.gt_conv(*a, plan)
    ref.backward(g)

    qb, eb = q.clone().requires_grad_(True), e.clone().requires_grad_(True)
    ko, vo = (t[:n_own].clone().requires_grad_(True) for t in (k, v))
    kh, vh = (t[n_own:].clone().requires_grad_(True) for t in (k, v))
    out = ops.gt_conv(qb, ko, vo, eb, plan, halo=(kh, vh))
    out.backward(g)
    torch.cuda.synchronize()

    assert rel_err(out, ref) <= tol
    assert rel_err(qb.grad, a[0].grad) <= tol and rel_err(eb.grad, a[3].grad) <= tol
    assert rel_err(torch.cat([ko.grad, kh.grad]), a[1].grad) <= tol
    assert rel_err(torch.cat([vo.grad, vh.grad]), a[2].grad) <= tol
