"""ROUND-2 work in progress (csrc/gtconv_fold.cu: GraphTransformerConv with lin_edge folded in) -- first contact with a GPU.

These kernels were written when the round's GPU budget was nearly spent: they have run exactly once (profiles/r01/fold_draft_r01ah.log,
all four cases green).  They are not on the default path (nothing calls them unless AB2_EDGE_FOLD=1).  This test runs them in a
CHILD process with a timeout, so that whatever they do cannot touch the CUDA context of the parity suite, and reports a failure as
`xfail`: the suite's colour is about the product path.  A pass means: folded conv == conv(q, k, v, lin_edge(raw)) of the product path,
outputs and every gradient (fp32 <= 1e-5, bf16 <= 2e-2 of max|ref|), on four head layouts."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, time, torch
sys.path.insert(0, %(root)r)
import anemoi_models_b200 as b2
from anemoi_models_b200 import ops
from anemoi_models_b200.graph import GraphCSR

dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)

def case(ns, nd, E, H, C, ed, dtype, tol, bias=True):
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, max(nd - 3, 1), (E,), generator=gen)]).to(dev)
    q, k, v = (torch.randn(n, H, C, generator=gen).to(dev, dtype) for n in (nd, ns, ns))
    raw = torch.rand(E, ed, generator=gen).to(dev)
    lin = torch.nn.Linear(ed, H * C, bias=bias).to(dev)
    g = torch.randn(nd, H, C, generator=gen).to(dev, dtype)
    plan = GraphCSR(ei, ns, nd)
    a = [t.clone().requires_grad_(True) for t in (q, k, v, raw)]
    ref = ops.gt_conv(a[0], a[1], a[2], lin(a[3]).to(dtype).view(E, H, C), plan)
    ref.backward(g)
    ref_g = [t.grad.float() for t in a] + [p.grad.clone().float() for p in lin.parameters()]
    lin.zero_grad()
    b = [t.clone().requires_grad_(True) for t in (q, k, v, raw)]
    out = ops.gt_conv_folded(b[0], b[1], b[2], b[3], lin.weight, lin.bias, plan)
    out.backward(g)
    got_g = [t.grad.float() for t in b] + [p.grad.float() for p in lin.parameters()]
    torch.cuda.synchronize()
    names = ["dq", "dk", "dv", "draw", "dW", "db"]
    worst = float((out.float() - ref.float()).abs().max() / max(1.0, float(ref.float().abs().max())))
    assert worst <= tol, ("out", worst)
    for n, x, y in zip(names, got_g, ref_g):
        err = float((x - y).abs().max() / max(1.0, float(y.abs().max())))
        assert err <= tol, (n, err, (ns, nd, E, H, C, str(dtype)))
        worst = max(worst, err)
    print("ok", (ns, nd, E, H, C, ed), dtype, "worst rel err %%.2e" %% worst, flush=True)

case(300, 120, 2000, 4, 8, 11, torch.float32, 1e-5)              # 2 lanes per head, 8 raw columns per lane
case(900, 400, 6000, 16, 16, 5, torch.float32, 1e-5, bias=False)  # 4 lanes per head
case(5000, 2000, 30000, 16, 64, 11, torch.bfloat16, 2e-2)        # the model's layout: 8 lanes per head, 2 columns per lane
case(700, 300, 4000, 8, 128, 15, torch.bfloat16, 2e-2)           # 16 lanes per head, 1 column per lane

# timing at the headline shape (edges per dst ~ 18.6): folded vs product path incl. its lin_edge GEMM
ns, nd, E, H, C, ed = 542080, 40320, 748256, 16, 64, 11
dst = torch.arange(E, device=dev) %% nd
src = torch.randint(0, ns, (E,), device=dev)
ei = torch.stack([src, dst.sort().values])
plan = GraphCSR(ei, ns, nd)
q, k, v = (torch.randn(n, H, C, device=dev, dtype=torch.bfloat16).requires_grad_(True) for n in (nd, ns, ns))
raw = torch.rand(E, ed, device=dev).requires_grad_(True)
lin = torch.nn.Linear(ed, H * C).to(dev)
g = torch.randn(nd, H, C, device=dev, dtype=torch.bfloat16)
def unfused():
    ops.gt_conv(q, k, v, lin(raw).to(torch.bfloat16).view(E, H, C), plan).backward(g)
def folded():
    ops.gt_conv_folded(q, k, v, raw, lin.weight, lin.bias, plan).backward(g)
for name, fn in (("lin_edge + conv (product path)", unfused), ("folded draft", folded)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        fn()
    t1.record(); torch.cuda.synchronize()
    print("%%-34s %%.3f ms per fwd+bwd" %% (name, t0.elapsed_time(t1) / 10), flush=True)

# block level: the reference blocks' golden vectors through the AB2_EDGE_FOLD branch (fp32, <= 1e-5 of max|ref|)
sys.path.insert(0, %(root)r + "/tests")
from conftest import load_golden, rel_err, t
from anemoi_models_b200.layers import block as b2block
assert b2block._EDGE_FOLD, "AB2_EDGE_FOLD=1 must be set before the import"
for fixture, kind in (("block_gt_mapper.npz", "mapper"), ("block_gt_processor.npz", "processor")):
    z = load_golden(fixture)
    ns, nd, D, H, ed, hid = (int(x) for x in z["meta"])
    cls = b2.GraphTransformerMapperBlock if kind == "mapper" else b2.GraphTransformerProcessorBlock
    blk = cls(D, hid, D, edge_dim=ed, num_heads=H).to(dev)
    blk.load_state_dict({k_[2:]: t(v_) for k_, v_ in z.items() if k_.startswith("p.")})
    ei, ea = t(z["edge_index"]).to(dev), t(z["ea"]).to(dev).requires_grad_(True)
    launches0 = ops._lib.lib().ab2_launch_count()
    if kind == "mapper":
        xs, xd = t(z["xs"]).to(dev).requires_grad_(True), t(z["xd"]).to(dev).requires_grad_(True)
        (_, out), _ = blk((xs, xd), ea, ei, ([[ns, D]], [[nd, D]], [[ea.shape[0], ed]]), 1, size=(ns, nd))
        ref_out, grads = t(z["dst_new"]), [(xs, "dxs"), (xd, "dxd"), (ea, "dea")]
    else:
        xd = t(z["x"]).to(dev).requires_grad_(True)
        out, _ = blk(xd, ea, ei, ([[nd, D]], [[nd, D]], [[ea.shape[0], ed]]), 1)
        ref_out, grads = t(z["nodes_new"]), [(xd, "dx"), (ea, "dea")]
    (out * t(z["gd"]).to(dev)).sum().backward()
    assert rel_err(out, ref_out) < 1e-5, ("out", kind, rel_err(out, ref_out))
    for x_, key in grads:
        assert rel_err(x_.grad, t(z[key])) < 1e-5, (key, kind, rel_err(x_.grad, t(z[key])))
    for name, p_ in blk.named_parameters():
        assert rel_err(p_.grad, t(z["gp." + name])) < 2e-5, (name, kind)
    print("ok block", kind, "through the folded branch; library launches:", ops._lib.lib().ab2_launch_count() - launches0, flush=True)
print("FOLD_DRAFT_OK")
"""


def test_folded_draft_kernels_in_a_child_process():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    env = dict(os.environ, AB2_EDGE_FOLD="1")
    try:
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
        report = r.stdout[-3000:] + "\n" + r.stderr[-3000:]
        ok = r.returncode == 0 and "FOLD_DRAFT_OK" in r.stdout
    except subprocess.TimeoutExpired as ex:
        report, ok = f"timeout: {ex}", False
    try:  # keep the report where a gpurun call brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "fold_draft_report.txt"), "w") as f:
            f.write(report)
    except OSError:
        pass
    print(report)
    if not ok:
        pytest.xfail("round-2 draft kernels (not on the product path) did not pass their first GPU run:\n" + report[-1500:])
