"""Timing breakdown of the halo exchange (run under torchrun): which part of gt_conv_sharded's exchange costs what."""
import os, sys, json, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from anemoi_models_b200.distributed.halo import build_local_halo_plan, exchange_rows, return_rows

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    ei_np, Ns, Nd, sb, db = bench.build_shard(world, rank)
    ei = torch.from_numpy(ei_np).to(dev)
    hp = build_local_halo_plan(ei, sb, db, group)
    H, C = 16, 64
    k = torch.randn(hp.n_own, H, C, device=dev, dtype=torch.bfloat16)
    dk = torch.zeros_like(k)
    gh = torch.randn(hp.n_halo, H, C, device=dev, dtype=torch.bfloat16)
    res = {"world": world, "n_halo": hp.n_halo, "send": sum(hp.send_counts), "send_counts": hp.send_counts, "recv_counts": hp.recv_counts}
    res["index_select_ms"] = timeit(lambda: k.index_select(0, hp.send_idx))
    send = k.index_select(0, hp.send_idx).contiguous()
    recv = torch.empty((hp.n_halo, H, C), device=dev, dtype=torch.bfloat16)
    res["a2a_fwd_ms"] = timeit(lambda: dist.all_to_all_single(recv, send, hp.recv_counts, hp.send_counts, group=group))
    back = torch.empty_like(send)
    res["a2a_bwd_ms"] = timeit(lambda: dist.all_to_all_single(back, gh, hp.send_counts, hp.recv_counts, group=group))
    def addall():
        off = 0
        for cnt in hp.send_counts:
            if cnt: dk.index_add_(0, hp.send_idx[off:off + cnt], back[off:off + cnt])
            off += cnt
    res["index_add_ms"] = timeit(addall)
    res["exchange_rows_ms"] = timeit(lambda: exchange_rows(k, hp, group))
    res["return_rows_ms"] = timeit(lambda: return_rows(gh, hp, group, dk))
    tiny = torch.zeros(1, device=dev)
    res["tiny_allreduce_ms"] = timeit(lambda: dist.all_reduce(tiny))
    mb = hp.n_halo * H * C * 2 / 1e6
    res["halo_MB_per_tensor"] = mb
    if rank == 0: print(json.dumps(res), flush=True)
    out = [None] * world
    dist.all_gather_object(out, {"rank": rank, "n_halo": hp.n_halo, "send": sum(hp.send_counts)})
    if rank == 0: print(json.dumps(out), flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
