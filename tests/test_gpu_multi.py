"""Multi-GPU parity over NCCL (needs >= 2 CUDA devices; skipped on a single-GPU box).

The dst-row-sharded blocks (halo exchange instead of the reference's head all-to-all / node all-gather) must give,
on P GPUs: (1) concatenated outputs == the single-GPU block output, (2) sum over ranks of weight grads == single-GPU
weight grads, (3) grads of the sharded inputs == slices of the single-GPU grads.  The single-GPU block is itself
pinned to the reference by tests/test_gpu_graphconv_blocks.py.
"""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, halo):
    try:
        os.environ["AB2_HALO"] = halo
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        blocks = _sharded_blocks(rank, world)
        plans = [entry[2] for b in blocks for entry in b.__dict__.get("_b200_halo_plans", {}).values()]
        plans.append(_sharded_conv_2kb_rows(rank, world))
        if halo == "p2p":  # the peer-memory transport must really have been used (no silent NCCL fallback)
            peers = [px for pl in plans for px in getattr(pl, "_peer", {}).values()]
            assert peers and all(px is not None and px.fwd_epoch > 0 and px.bwd_epoch > 0 for px in peers), \
                "NVLink peer-memory halo exchange was not used"
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None))
    except Exception:
        q.put((rank, traceback.format_exc()))


def _graph(ns, nd, deg, gen):
    """latitude-like locality: dst i draws `deg` src from a window around its own position."""
    base = (torch.arange(nd) * ns // nd).view(-1, 1)
    src = (base + torch.randint(-ns // 16, ns // 16, (nd, deg), generator=gen)).clamp_(0, ns - 1).view(-1)
    dst = torch.arange(nd).repeat_interleave(deg)
    perm = torch.randperm(src.numel(), generator=gen)  # shuffled edge order: perm != identity
    return torch.stack([src, dst])[:, perm].contiguous()


def _sharded_blocks(rank, world):
    import anemoi_models_b200 as b2
    from anemoi_models_b200.distributed.khop_edges import edge_chunk_order
    from anemoi_models_b200.distributed.shapes import bounds_from_shapes, get_shape_shards

    dev = torch.device("cuda", rank)
    group = dist.group.WORLD
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    ns, nd, D, H, ed, hid = 3001, 1203, 128, 8, 7, 256
    ei = _graph(ns, nd, 9, gen).to(dev)
    E = ei.shape[1]
    xs_f, xd_f = torch.randn(ns, D, generator=gen).to(dev), torch.randn(nd, D, generator=gen).to(dev)
    ea_f, gd_f = torch.rand(E, ed, generator=gen).to(dev), torch.randn(nd, D, generator=gen).to(dev)
    sh_s, sh_d, sh_e = (get_shape_shards(x, 0, group) for x in (xs_f, xd_f, ea_f))
    sb, db, eb = bounds_from_shapes(sh_s), bounds_from_shapes(sh_d), bounds_from_shapes(sh_e)

    # ---- GT mapper block
    blk = b2.GraphTransformerMapperBlock(D, hid, D, edge_dim=ed, num_heads=H).to(dev)  # same seed -> same weights on all ranks
    xs1, xd1, ea1 = (x.clone().requires_grad_(True) for x in (xs_f, xd_f, ea_f))
    (_, ref), _ = blk((xs1, xd1), ea1, ei, ([[ns, D]], [[nd, D]], [[E, ed]]), 1, size=(ns, nd))
    (ref * gd_f).sum().backward()
    ref_grads = {n: p.grad.clone() for n, p in blk.named_parameters()}
    blk.zero_grad()
    xs = xs_f[sb[rank]:sb[rank + 1]].clone().requires_grad_(True)
    xd = xd_f[db[rank]:db[rank + 1]].clone().requires_grad_(True)
    ea = ea_f[eb[rank]:eb[rank + 1]].clone().requires_grad_(True)
    (_, out), _ = blk((xs, xd), ea, ei, (sh_s, sh_d, sh_e), 1, group, size=(ns, nd))
    (out * gd_f[db[rank]:db[rank + 1]]).sum().backward()
    tol = 2e-5
    assert float((out - ref[db[rank]:db[rank + 1]]).abs().max()) < tol * max(1.0, float(ref.abs().max()))
    for got, full, b in ((xs.grad, xs1.grad, sb), (xd.grad, xd1.grad, db), (ea.grad, ea1.grad, eb)):
        want = full[b[rank]:b[rank + 1]]
        assert float((got - want).abs().max()) < tol * max(1.0, float(full.abs().max()))
    for n, p in blk.named_parameters():
        gsum = p.grad.clone()
        dist.all_reduce(gsum)
        assert float((gsum - ref_grads[n]).abs().max()) < 5e-5 * max(1.0, float(ref_grads[n].abs().max())), n

    # ---- GraphConv processor block (edges pre-partitioned by dst owner, as the reference's GNNProcessor does)
    n = nd
    ei2 = _graph(n, n, 6, gen).to(dev)
    E2 = ei2.shape[1]
    x_f, e_f = torch.randn(n, D, generator=gen).to(dev), torch.randn(E2, D, generator=gen).to(dev)
    gx_f, ge_f = torch.randn(n, D, generator=gen).to(dev), torch.randn(E2, D, generator=gen).to(dev)
    gblk = b2.GraphConvProcessorBlock(D, D).to(dev)
    x1, e1 = x_f.clone().requires_grad_(True), e_f.clone().requires_grad_(True)
    ref_n, ref_e = gblk(x1, e1, ei2, ([[n, D]], [[n, D]], None))
    ((ref_n * gx_f).sum() + (ref_e * ge_f).sum()).backward()
    ref_grads = {k: p.grad.clone() for k, p in gblk.named_parameters()}
    gblk.zero_grad()
    order, counts = edge_chunk_order(n, ei2, world)
    ids = torch.split(order, counts)[rank]
    sh = get_shape_shards(x_f, 0, group)
    nb = bounds_from_shapes(sh)
    x = x_f[nb[rank]:nb[rank + 1]].clone().requires_grad_(True)
    e = e_f[ids].clone().requires_grad_(True)
    nodes, edges = gblk(x, e, ei2[:, ids], (sh, sh, None), group)
    ((nodes * gx_f[nb[rank]:nb[rank + 1]]).sum() + (edges * ge_f[ids]).sum()).backward()
    assert float((nodes - ref_n[nb[rank]:nb[rank + 1]]).abs().max()) < tol * max(1.0, float(ref_n.abs().max()))
    assert float((edges - ref_e[ids]).abs().max()) < tol * max(1.0, float(ref_e.abs().max()))
    assert float((x.grad - x1.grad[nb[rank]:nb[rank + 1]]).abs().max()) < tol * max(1.0, float(x1.grad.abs().max()))
    assert float((e.grad - e1.grad[ids]).abs().max()) < tol * max(1.0, float(e1.grad.abs().max()))
    for k, p in gblk.named_parameters():
        gsum = p.grad.clone()
        dist.all_reduce(gsum)
        assert float((gsum - ref_grads[k]).abs().max()) < 5e-5 * max(1.0, float(ref_grads[k].abs().max())), k
    return [blk, gblk]


def _sharded_conv_2kb_rows(rank, world):
    """D = 1024 bf16 (2 KB rows: the bulk-copy pipelined kernels with a halo buffer, interior / boundary split and the peer
    push hidden behind the interior rows) with UNEVEN, dst-aligned src shards -- the configuration of the scaling benchmark --
    against the single-rank conv of the whole graph (every rank computes it; sizes are small)."""
    from anemoi_models_b200 import ops
    from anemoi_models_b200.distributed.halo import aligned_src_bounds, build_local_halo_plan
    from anemoi_models_b200.distributed.khop_edges import edge_chunk_order
    from anemoi_models_b200.distributed.shapes import tensor_split_sizes
    from anemoi_models_b200.graph import GraphCSR
    from conftest import rel_err, rel_l2

    dev = torch.device("cuda", rank)
    group = dist.group.WORLD
    gen = torch.Generator().manual_seed(7)
    ns, nd, H, C = 20011, 6007, 16, 64
    ei = _graph(ns, nd, 9, gen).to(dev)
    E = ei.shape[1]
    bf = torch.bfloat16
    q, g = (torch.randn(nd, H, C, generator=gen).to(dev, bf) for _ in range(2))
    k, v = (torch.randn(ns, H, C, generator=gen).to(dev, bf) for _ in range(2))
    e = torch.randn(E, H, C, generator=gen).to(dev, bf)
    full = [t.clone().requires_grad_(True) for t in (q, k, v, e)]
    ref = ops.gt_conv(*full, GraphCSR(ei, ns, nd))
    ref.backward(g)
    db = [0]
    for s_ in tensor_split_sizes(nd, world):
        db.append(db[-1] + s_)
    order, counts = edge_chunk_order(nd, ei, world)
    ids = torch.split(order, counts)[rank]
    ei_loc_glob = ei[:, ids].contiguous()
    sb = aligned_src_bounds(ei_loc_glob, ns, group)  # uneven src shards that follow the dst shards
    assert len(set(sb[r + 1] - sb[r] for r in range(world))) > 1 or world == 1
    hplan = build_local_halo_plan(ei_loc_glob, sb, db, group)
    plan = GraphCSR(hplan.local_edge_index, hplan.n_src, db[rank + 1] - db[rank])
    mine = [q[db[rank]:db[rank + 1]], k[sb[rank]:sb[rank + 1]], v[sb[rank]:sb[rank + 1]], e[ids]]
    mine = [t.clone().requires_grad_(True) for t in mine]
    out = ops.gt_conv_sharded(*mine, plan, hplan, group)
    out.backward(g[db[rank]:db[rank + 1]])
    want = [ref[db[rank]:db[rank + 1]], full[0].grad[db[rank]:db[rank + 1]], full[1].grad[sb[rank]:sb[rank + 1]],
            full[2].grad[sb[rank]:sb[rank + 1]], full[3].grad[ids]]
    got = [out, mine[0].grad, mine[1].grad, mine[2].grad, mine[3].grad]
    for name, a, b in zip(("out", "dq", "dk", "dv", "de"), got, want):
        assert rel_err(a, b) < 2e-2 and rel_l2(a, b) < 1e-2, (name, rel_err(a, b), rel_l2(a, b))
    return hplan


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("halo", ["p2p", "nccl"])  # NVLink peer-memory push (default) and the NCCL all-to-all
def test_sharded_blocks_nccl(world, halo):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, halo)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    errs = [f"rank {r}:\n{e}" for r, e in results if e]
    assert not errs, "\n".join(errs)
