"""debug runner of tests/test_gpu_cuda_graph.py's child script (PYTHONPATH=. python scripts/cuda_graph_check.py)"""

import torch
import anemoi_models_b200 as b2
from anemoi_models_b200 import synthetic as S

dev = torch.device("cuda", 0)
torch.manual_seed(0)
xyz, ei_np = S.multiscale_icosahedral_mesh(3)
N, E = xyz.shape[0], ei_np.shape[1]
ei = torch.from_numpy(ei_np).to(dev)
D = 256
for kind in ("gt", "graphconv"):
    if kind == "gt":
        blk = b2.GraphTransformerProcessorBlock(D, 4 * D, D, edge_dim=11, num_heads=8).to(dev).to(torch.bfloat16)
        e = torch.randn(E, 11, device=dev, dtype=torch.bfloat16)
        call = lambda xx, ee: blk(xx, ee, ei, (None, None, None), 1)[0]
    else:
        blk = b2.GraphConvProcessorBlock(D, D).to(dev).to(torch.bfloat16)
        e = torch.randn(E, D, device=dev, dtype=torch.bfloat16)
        call = lambda xx, ee: blk(xx, ee, ei, ([[N, D]], [[N, D]], None))[0]
    x = torch.randn(N, D, device=dev, dtype=torch.bfloat16)
    g = torch.randn(N, D, device=dev, dtype=torch.bfloat16)
    xs, es = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    side = torch.cuda.Stream()  # the PyTorch recipe: eager warm-up on a side stream, nothing of its autograd graph kept alive
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):  # plans, shared-memory attributes, allocator
            call(xs, es).backward(g)
        xs.grad = es.grad = None
        for p in blk.parameters():
            p.grad = None
        out_e = call(xs, es)
        out_e.backward(g)
        ref = [out_e.detach().clone(), xs.grad.clone(), es.grad.clone()] + [p.grad.clone() for p in blk.parameters()]
        del out_e
        xs.grad = es.grad = None
        for p in blk.parameters():
            p.grad = None
    torch.cuda.current_stream().wait_stream(side)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        out_g = call(xs, es)
        out_g.backward(g)
    for _ in range(2):
        cg.replay()
    torch.cuda.synchronize()
    got = [out_g, xs.grad, es.grad] + [p.grad for p in blk.parameters()]
    assert len(got) == len(ref)
    for i, (a, b) in enumerate(zip(got, ref)):
        assert torch.equal(a, b), (kind, i, float((a.float() - b.float()).abs().max()))
print("ok")
