#!/usr/bin/env python
"""bench.py -- GT-conv edges/s, forward+backward, on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU op sequence (oracle port) on host cores

A "step" = one GraphTransformerConv forward + backward (conv boundary, reference layers/conv.py:98) over one
synthetic graph.  N=1: BASELINE configs[1] -- GT mapper encoder n320 (542,080 pts, Fibonacci sphere) -> o96 (40,320),
hidden 1024, 16 heads, bf16, cut-off-0.6 edges (E = 748,256).  N>1 (weak scaling, one process per GPU): the same graph
family at N x the node counts (Fibonacci N*542,080 -> o<N'> with ~N*40,320 points), dst rows sharded by tensor_split
like the model's shard shapes, each rank owning the edges into its rows; every step exchanges the halo of k/v rows
(NCCL all-to-all over NVLink) before the forward and returns the halo gradients after the backward.

One JSON line on stdout (rank 0).  `value` = edges/s with inputs resident in HBM; `e2e` = the same step through the
host-buffer C-ABI entry point (pinned host tensors in, H2D/D2H copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, C = 16, 64
D = H * C
SRC_POINTS, DST_N = 542080, 96
METRIC, UNIT = "gt_conv_fwd_bwd_edges_per_s", "edges/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers allocated afterwards
    (first touch by this process) live in the memory next to the GPU's PCIe root port.  Round 1 allocated every rank's buffers
    wherever the launcher happened to run (all ranks on node 0): the host-buffer arm anti-scaled (0.20 at 8 GPUs).
    Returns a small dict for the bench line; does nothing (and says so) when sysfs does not expose the topology."""
    info = {"numa_node": None, "cpus": None}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["pci"] = bdf
        if node < 0:
            info["note"] = "sysfs reports no NUMA node for the GPU (single-node host or virtualised topology)"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        info.update(numa_node=node, cpus=len(use))
    except Exception as ex:  # noqa: BLE001
        info["note"] = f"not bound: {type(ex).__name__}: {ex}"[:200]
    return info


def dst_N_for(parts: int) -> int:
    """octahedral resolution whose point count is closest to parts * 40,320 (o96 for parts = 1)."""
    target = parts * (4 * DST_N * DST_N + 36 * DST_N)
    n = int(round((-36 + (36 * 36 + 16 * target) ** 0.5) / 8))
    return n


def workload_name(parts: int, workload: str = "encoder") -> str:
    if workload.startswith("config1"):
        return f"BASELINE configs[0] conv shape {workload}: o96 (40,320) / o48 (10,944) grids, D=256, H=16, fp32 fwd+bwd (report line; launch-bound at this size)"
    if workload == "decoder":
        return "GT mapper decoder conv o96(40320) -> n320(542080, Fibonacci), 3-NN, D=1024, H=16, bf16 fwd+bwd (report line, not the headline)"
    if workload == "processor":
        return "GT processor conv on o96(40320), 8-NN, D=1024, H=16, bf16 fwd+bwd (report line, not the headline)"
    if parts == 1:
        return "GT mapper encoder conv n320(542080, Fibonacci) -> o96(40320), cut-off 0.6, D=1024, H=16, bf16 fwd+bwd"
    return (f"same graph family at {parts}x area: Fibonacci({parts * SRC_POINTS}) -> o{dst_N_for(parts)}, dst-row sharded over "
            f"{parts} ranks with k/v halo exchange, D=1024, H=16, bf16 fwd+bwd")


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # NVML missing: report that instead of inventing numbers
            self.nv, self.err = None, repr(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def build_shard(parts: int, part: int, workload: str = "encoder", weak_split: str = "equal"):
    """edge_index (GLOBAL ids) of the edges into shard `part`, plus global sizes and shard bounds."""
    from anemoi_models_b200 import synthetic as S
    from anemoi_models_b200.distributed.shapes import tensor_split_sizes

    ns = parts * SRC_POINTS
    if workload == "encoder":
        bounds = None
        if parts > 1 and weak_split == "work":  # equal WORK per rank (what weak scaling means), not equal dst-row counts
            bounds = S.encoder_work_balanced_bounds(ns, dst_N_for(parts), parts)
        ei, ns, nd, radius = S.encoder_graph_band(ns, dst_N_for(parts), parts, part, bounds=bounds)
        if bounds is not None:
            sb = np.concatenate([[0], np.cumsum(tensor_split_sizes(ns, parts))]).tolist()
            return ei, ns, nd, sb, [int(b) for b in bounds]
    else:
        assert parts == 1, "decoder / processor workloads are single-GPU report lines"
        if workload.startswith("config1"):  # BASELINE configs[0]: o96 data grid, o48 hidden grid
            data, _ = S.octahedral_grid(96)
            hid48, _ = S.octahedral_grid(48)
            if workload == "config1-enc":
                ei, ns, nd = S.cutoff_edges(data, hid48, 0.6 * S.max_nn_distance(hid48)), len(data), len(hid48)
            elif workload == "config1-proc":
                ei, ns, nd = S.knn_edges(hid48, hid48, 8, exclude_self=True), len(hid48), len(hid48)
            else:
                ei, ns, nd = S.knn_edges(hid48, data, 3), len(hid48), len(data)
            sb = [0, ns]
            db = [0, nd]
            return ei, ns, nd, sb, db
        hidden, _ = S.octahedral_grid(DST_N)
        if workload == "decoder":  # o96 -> n320, 3 nearest hidden nodes per data node
            ei, ns, nd = S.knn_edges(hidden, S.fibonacci_sphere(SRC_POINTS), 3), len(hidden), SRC_POINTS
        else:  # processor on o96: 8 nearest neighbours, no self loops
            ei, ns, nd = S.knn_edges(hidden, hidden, 8, exclude_self=True), len(hidden), len(hidden)
    sb = np.concatenate([[0], np.cumsum(tensor_split_sizes(ns, parts))]).tolist()
    db = np.concatenate([[0], np.cumsum(tensor_split_sizes(nd, parts))]).tolist()
    return ei, ns, nd, sb, db


def algorithmic_bytes(E, Ns, Nd, b=2, D=D, H=H):
    """Compulsory HBM traffic, each tensor touched once per kernel (DESIGN.md 'roofline accounting')."""
    idx_fwd = 4 * (2 * E + Nd + 1)  # col + perm + rowptr
    fwd = b * (E * D + 2 * Ns * D + 2 * Nd * D) + idx_fwd + 4 * Nd * H
    bwd_dst = b * (2 * E * D + 2 * Ns * D + 4 * Nd * D) + idx_fwd + 4 * Nd * H + 8 * E * H
    bwd_src = b * (2 * Ns * D + 2 * Nd * D) + 8 * E * H + 4 * (2 * E + Ns + 1)
    # SURVEY 8d / BASELINE.md step figure (the north_star target is quoted on this one; it does not count the small
    # [E,H] softmax-weight workspace or the second read of q and g by the src pass)
    step = b * (3 * E * D + 6 * Ns * D + 6 * Nd * D) + 8 * (E + Nd + 1) + 8 * Nd * H
    return {"fwd": fwd, "bwd_dst": bwd_dst, "bwd_src": bwd_src, "step_survey": step}


class _SkipE2E(Exception):
    pass


def run_ours(args):
    import torch.distributed as dist

    import anemoi_models_b200 as b2
    from anemoi_models_b200 import _lib, ops
    from anemoi_models_b200.distributed.halo import build_local_halo_plan
    from anemoi_models_b200.graph import GraphCSR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    L = _lib.lib()
    cfg1 = args.workload.startswith("config1")
    H, C = (16, 16) if cfg1 else (16, 64)
    D = H * C
    dt_code, esz = (0, 4) if cfg1 else (1, 2)
    ei_np, Ns_g, Nd_g, sb, db = build_shard(world, rank, args.workload, args.weak_split)
    ei_glob = torch.from_numpy(ei_np).to(dev)
    E = ei_glob.shape[1]
    torch.manual_seed(1234 + rank)
    bf = torch.float32 if cfg1 else torch.bfloat16
    if world > 1 and args.src_split == "aligned":
        # src ownership follows dst ownership (caller-chosen shard shapes; see distributed/halo.py aligned_bounds_from_ranges):
        # equal-count src shards of the uniform Fibonacci grid do not align in latitude with equal-count dst shards of the
        # octahedral grid, which makes 9-25 % of a shard halo at 4-8 ranks
        from anemoi_models_b200.distributed.halo import aligned_src_bounds

        sb = aligned_src_bounds(ei_glob, Ns_g, group)
    nd_loc, ns_loc = db[rank + 1] - db[rank], sb[rank + 1] - sb[rank]
    q = torch.randn(nd_loc, H, C, device=dev, dtype=bf)
    g = torch.randn(nd_loc, H, C, device=dev, dtype=bf)
    e = torch.randn(E, H, C, device=dev, dtype=bf)
    k_own = torch.randn(ns_loc, H, C, device=dev, dtype=bf)
    v_own = torch.randn(ns_loc, H, C, device=dev, dtype=bf)

    if world > 1:
        hplan = build_local_halo_plan(ei_glob, sb, db, group)
        ei_loc = hplan.local_edge_index
        n_src = hplan.n_src
    else:
        hplan, ei_loc, n_src = None, ei_glob, Ns_g
    plan = GraphCSR(ei_loc, n_src, nd_loc)
    conv = b2.GraphTransformerConv(out_channels=C)

    def step():
        """one forward + backward of the conv boundary; returns nothing (grads land in .grad buffers)"""
        qq, ee = q.detach().requires_grad_(True), e.detach().requires_grad_(True)
        kk, vv = k_own.detach().requires_grad_(True), v_own.detach().requires_grad_(True)
        if world > 1:
            out = ops.gt_conv_sharded(qq, kk, vv, ee, plan, hplan, group)
        else:
            out = conv(qq, kk, vv, ee, ei_loc, (n_src, nd_loc), plan=plan)
        out.backward(g)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # multi-GPU: before anything is timed, the sharded step is checked against a single-rank conv of the same sub-graph
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = sharded_parity(q, k_own, v_own, e, g, ei_glob, Ns_g, sb, rank, world, plan, hplan, group, dev)
        parity["within_tolerance"] = bool(parity["max_rel_err"] <= 2e-2 and parity["max_rel_l2"] <= 2e-2)
        if not parity["within_tolerance"] and rank == 0:  # reported in the line (and loudly here); the timing below is then void
            print(f"[bench] PARITY FAILURE: sharded conv differs from the single-rank conv: {parity}", file=sys.stderr, flush=True)
    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        barrier()
        launches0 = L.ab2_launch_count()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    launches = L.ab2_launch_count() - launches0  # kernels of libanemoi_b200 launched by this rank inside the timed region
    ms = ev0.elapsed_time(ev1)
    if os.environ.get("AB2_TRACE", "0") == "1" and world > 1:  # debug: phase times of one step on every rank
        ops.trace_summary()
        barrier()
        step()
        torch.cuda.synchronize()
        ph = ops.trace_summary()
        allph = [None] * world
        dist.all_gather_object(allph, [(lab, round(t, 3)) for lab, t in ph])
        if rank == 0:
            for r, p_ in enumerate(allph):
                print(f"[trace rank {r}] " + "  ".join(f"{lab}={t}" for lab, t in p_), file=sys.stderr, flush=True)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    etot = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    shard_stats = torch.tensor([float(ns_loc), float(hplan.n_halo if hplan is not None else 0), float(E), float(nd_loc)],
                               device=dev, dtype=torch.float64)
    shard_max, shard_min = shard_stats.clone(), shard_stats.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(etot, op=dist.ReduceOp.SUM)
        dist.all_reduce(shard_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(shard_min, op=dist.ReduceOp.MIN)
    ms_per_step = float(tmax) / args.steps
    value = float(etot) / (ms_per_step * 1e-3)

    # ---- the same step on the reference's equal-count dst shards (`tensor_split`), for the record
    other_split = None
    if world > 1 and args.workload == "encoder" and args.weak_split == "work":
        try:
            other_split = sharded_step_other_split(args, world, rank, dev, group, "equal", H, C, bf)
        except Exception as ex:  # noqa: BLE001 -- the headline line must still be printed
            other_split = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- per-kernel durations (CUDA events on the launching stream, rank 0's shard) -> roofline of the dominant kernel
    st = torch.cuda.current_stream(dev).cuda_stream
    out = torch.empty_like(q)
    lse2 = torch.empty(nd_loc, H, device=dev, dtype=torch.float32)
    kn = torch.randn(n_src, H, C, device=dev, dtype=bf) if world > 1 else k_own
    vn = torch.randn(n_src, H, C, device=dev, dtype=bf) if world > 1 else v_own
    dq, de, dk, dv = torch.empty_like(q), torch.empty_like(e), torch.empty_like(kn), torch.empty_like(vn)
    ws_bytes = L.ab2_gtconv_bwd_workspace_bytes(E, H)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    P = _lib.ptr
    kreps = 20
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(kreps)]

    def kernels(ev=None):
        if ev: ev[0].record()
        _lib.check(L.ab2_gtconv_fwd(P(q), P(kn), P(vn), P(e), dt_code, P(plan.rowptr), P(plan.col), P(plan.perm), n_src, nd_loc, E, H, C,
                                    P(out), P(lse2), st))
        if ev: ev[1].record()
        _lib.check(L.ab2_gtconv_bwd_dst(P(q), P(kn), P(vn), P(e), dt_code, P(plan.rowptr), P(plan.col), P(plan.perm), P(plan.csr2csc),
                                        n_src, nd_loc, E, H, C, P(out), P(lse2), P(g), P(dq), P(de), P(ws), ws_bytes, st))
        if ev: ev[2].record()
        _lib.check(L.ab2_gtconv_bwd_src(P(q), P(g), dt_code, P(plan.colptr), P(plan.crow), n_src, nd_loc, E, H, C, P(ws), P(dk), P(dv), st))
        if ev: ev[3].record()

    for _ in range(3):
        kernels()
    torch.cuda.synchronize()
    with sampler:
        for r in range(kreps):
            kernels(evs[r])
        torch.cuda.synchronize()
    kt = {name: float(np.mean([evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(kreps)]))
          for i, name in enumerate(("fwd", "bwd_dst", "bwd_src"))}
    ab = algorithmic_bytes(E, n_src, nd_loc, esz, D, H)
    peak, peak_src = peaks()
    dom = max(kt, key=kt.get)
    achieved = ab[dom] / (kt[dom] * 1e-3) / 1e9
    traffic = None  # dram__bytes_read+write per launch from the committed `ncu --set full` capture of this same workload
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world == 1 and args.workload == "encoder" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(L.ab2_gtconv_variant({"fwd": 0, "bwd_dst": 1, "bwd_src": 2}[dom], dt_code, n_src, nd_loc, E, H, C).decode().split("<")[0],
                                       {}).get("dram_bytes_per_launch")
    which = {"fwd": 0, "bwd_dst": 1, "bwd_src": 2}
    kname = {n: L.ab2_gtconv_variant(which[n], dt_code, n_src, nd_loc, E, H, C).decode() for n in kt}
    roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab[dom], "launch_ms": round(kt[dom], 4),
                "note": "peak = measured COPY bandwidth (half reads, half writes); a read-dominated kernel can exceed it (frac > 1): "
                        "ncu reports the same kernel at ~84 % of the 7.7 TB/s HBM3e pin rate, and traffic ~= algorithmic bytes "
                        "(profiles/r01/ncu_gtconv_r01ab.md); write-only traffic on this GPU tops out at ~3.9 TB/s (hbm_probe_r01ab.json)"}
    kern = {n: {"kernel": kname[n], "ms": round(kt[n], 4), "algorithmic_GB": round(ab[n] / 1e9, 3), "GBps": round(ab[n] / (kt[n] * 1e-3) / 1e9, 1),
                "frac_of_peak": round(ab[n] / (kt[n] * 1e-3) / 1e9 / peak, 4)} for n in kt}
    t_kernels = sum(kt.values())
    step_gbps = ab["step_survey"] / (t_kernels * 1e-3) / 1e9

    # ---- e2e: the same step through the host-buffer C-ABI call (pinned host tensors in and out), rank-local
    e2e = None
    host = outs = dev_ws = None
    err = ""
    numa = bind_to_gpu_numa(local_rank)  # before the pinned buffers are allocated
    try:
        if args.e2e_steps <= 0:
            raise _SkipE2E()
        # N > 1: every rank feeds ITS shard (compact src space [own | halo]: the host supplies the halo rows' k, v and gets
        # their dk, dv back) through its own PCIe link, all ranks at once; time = max over ranks
        if world > 1:
            k_h = torch.cat([k_own, torch.randn(n_src - ns_loc, H, C, device=dev, dtype=bf)])
            v_h = torch.cat([v_own, torch.randn(n_src - ns_loc, H, C, device=dev, dtype=bf)])
        else:
            k_h, v_h = k_own, v_own
        host = [x.cpu().pin_memory() for x in (q, k_h, v_h, e, g)]
        del k_h, v_h
        need = L.ab2_gtconv_host_workspace_bytes(n_src, nd_loc, E, H, C, dt_code)
        dev_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        outs = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in (host[0], host[0], host[1], host[2], host[3])]
    except _SkipE2E:
        err = "skipped"
    except Exception as ex:  # e.g. no pinned host memory for N shards on this box
        if world == 1:
            raise
        err = f"{type(ex).__name__}: {ex}"[:300]
    ok = torch.tensor([0.0 if err else 1.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks take the same branch (no rank waits in a collective alone)
    if args.e2e_steps <= 0:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "skipped (--e2e-steps 0)"}
    elif float(ok) < 1.0:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "error": err or "host buffers could not be allocated on another rank"}
    else:
        def host_step():
            ops.gt_conv_host(*host, plan, dev_ws=dev_ws, outs=outs, nchunks=args.e2e_chunks)

        host_step()
        barrier()
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        with sampler:
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                host_step()  # synchronises internally: results are in host memory when it returns
            t1 = time.perf_counter()
        e2e_t = torch.tensor([(t1 - t0) * 1e3 / n_e2e], device=dev, dtype=torch.float64)
        io = torch.tensor([float(sum(x.numel() * x.element_size() for x in host)),
                           float(sum(x.numel() * x.element_size() for x in outs))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(io, op=dist.ReduceOp.SUM)
        e2e_ms = float(e2e_t)
        e2e = {"value": float(etot) / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": round(e2e_ms, 3), "steps": n_e2e,
               "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
               "api": ("ab2_gtconv_fwd_bwd_host_streamed" if plan.perm_is_identity and args.e2e_chunks > 1 else "ab2_gtconv_fwd_bwd_host")
                      + " (pinned host q,k,v,e,g in; out,dq,dk,dv,de back to pinned host; copies inside the call"
                      + ("; one call per rank on its shard, all ranks concurrently, max over ranks)" if world > 1 else ")"),
               "chunks": args.e2e_chunks, "numa": numa,
               "GBps_per_direction_per_gpu": round(float(io[0]) / world / (e2e_ms * 1e-3) / 1e9, 1)}
    del host, outs, dev_ws

    # ---- BASELINE configs[4]: the o1280 -> n320 graph, whole graph over `world` ranks (strong scaling; extra block of the line)
    config5 = None
    if args.workload == "encoder" and args.config5 != "off":
        del q, g, e, k_own, v_own, kn, vn, dq, de, dk, dv, ws, out, lse2
        torch.cuda.empty_cache()
        try:
            config5 = {"equal": config5_block(world, rank, dev, group, steps=min(args.steps, 20), warmup=3, dst_split="equal",
                                              parity=not args.no_parity_check)}
            if world > 1:
                config5["balanced"] = config5_block(world, rank, dev, group, steps=min(args.steps, 20), warmup=3, dst_split="balanced",
                                                    parity=False)
        except Exception as ex:  # noqa: BLE001 -- the headline line must still be printed
            config5 = {"error": f"{type(ex).__name__}: {ex}"[:400]}

    # ---- BASELINE configs[3]: step time of the AIFS-like n320/o96 model on this repo's blocks (the second half of the metric)
    model_step = None
    if args.workload == "encoder" and world == 1 and not args.no_model_step:
        try:
            import copy

            margs = copy.copy(args)
            margs.steps, margs.warmup, margs.profile = 5, 3, False
            ml = run_model(margs, emit=False)
            model_step = {"ms_per_step": ml["ms_per_step"], "workload": ml["config"]["workload"], "clocks": ml["clocks"],
                          "peak_mem_GB": ml["peak_mem_GB"], "loss": ml["config"]["loss"],
                          "optimizer_steps_before_loss": ml["config"]["optimizer_steps_before_loss"]}
            try:  # the same step with every activation kept in HBM instead of recomputed (180 GB: no need to checkpoint)
                nl = run_model(margs, emit=False, recompute=False)
                model_step["without_activation_checkpointing"] = {"ms_per_step": nl["ms_per_step"], "peak_mem_GB": nl["peak_mem_GB"],
                                                                  "loss": nl["config"]["loss"]}
            except Exception as ex:  # noqa: BLE001
                model_step["without_activation_checkpointing"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                torch.cuda.empty_cache()
            if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "anemoi", "models")):
                try:  # the same model from the unmodified reference blocks, on the same GPU
                    margs.steps = 2
                    rl = run_model(margs, emit=False, impl="reference")
                    model_step["reference_blocks_same_gpu"] = {"ms_per_step": rl["ms_per_step"], "peak_mem_GB": rl["peak_mem_GB"],
                                                               "loss": rl["config"]["loss"],
                                                               "optimizer_steps_before_loss": rl["config"]["optimizer_steps_before_loss"],
                                                               "what": "the same model wired from the unmodified reference blocks (baseline/_ref, PyG op "
                                                                       "sequence as torch ops, nn.Linear / nn.LayerNorm), same GPU, same autocast"}
                except Exception as ex:  # noqa: BLE001
                    model_step["reference_blocks_same_gpu"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                    torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001 -- the headline line must still be printed
            model_step = {"error": f"{type(ex).__name__}: {ex}"[:400]}

    # ---- CPU baseline: the reference's op sequence (oracle port) on this box's host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "encoder":
        cpu_baseline = cpu_reference_sample(ei_np, Ns_g, Nd_g, steps=2, warmup=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if cfg1 else "bf16", "data": "synthetic",
            "config": {"workload": workload_name(world, args.workload), "edges_total": int(float(etot)), "edges_rank0": int(E),
                       "src_rows_rank0": int(n_src), "halo_rows_rank0": int(hplan.n_halo) if hplan is not None else 0, "dst_rows_rank0": int(nd_loc), "hidden": D, "heads": H,
                       "src_split": (args.src_split if world > 1 else "n/a"),
                       "dst_split": ("n/a" if world == 1 else
                                     "equal work per rank (measured weights: 1.7 x edges + src rows; synthetic.encoder_work_balanced_bounds)"
                                     if args.weak_split == "work" else "equal dst-row counts (tensor_split, the reference's shapes)"),
                       "dst_rows_min_max_rank": [int(shard_min[3]), int(shard_max[3])],
                       "own_src_rows_max_rank": int(shard_max[0]),
                       "halo_rows_max_rank": int(shard_max[1]), "edges_max_rank": int(shard_max[2]),
                       "l2": ("inputs (>5 GB per step) exceed the 126 MB L2; no flush between steps" if not cfg1 else
                              "config-1 working set is L2-sized: steady-state (warm L2) numbers, no flush"),
                       "timed_region": "conv forward + backward (+ halo all-to-all of k,v and its backward when n_gpus>1); CSR build excluded (one-off, cached)"},
            "clocks": sampler.summary(),
            "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                 "note": "host-buffer arm is measured at n_gpus=1"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_step": {"achieved": round(step_gbps, 1), "peak": peak, "unit": "GB/s", "frac": round(step_gbps / peak, 4),
                              "bytes": ab["step_survey"], "kernel_ms_sum": round(t_kernels, 4),
                              "note": "SURVEY 8d algorithmic bytes of fwd+bwd over the sum of the three kernel durations"},
            "kernels": kern,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "equal_count_dst_split": other_split,
            "config5_o1280_to_n320": config5,
            "model_step_n320_o96": model_step,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
# multi-GPU: driver-visible parity of the sharded step, and BASELINE configs[4] (o1280 -> n320, strong scaling)
# ----------------------------------------------------------------------------------------------------------------
def sharded_step_other_split(args, world, rank, dev, group, weak_split, H, C, bf):
    """The weak-scaling step of the headline on the OTHER dst split (sub-block of the line): same graph, same timing rules."""
    import torch.distributed as dist

    from anemoi_models_b200 import ops
    from anemoi_models_b200.distributed.halo import aligned_src_bounds, build_local_halo_plan
    from anemoi_models_b200.graph import GraphCSR

    ei_np, Ns_g, Nd_g, sb, db = build_shard(world, rank, args.workload, weak_split)
    ei_glob = torch.from_numpy(ei_np).to(dev)
    E = ei_glob.shape[1]
    if args.src_split == "aligned":
        sb = aligned_src_bounds(ei_glob, Ns_g, group)
    nd_loc, ns_loc = db[rank + 1] - db[rank], sb[rank + 1] - sb[rank]
    torch.manual_seed(4321 + rank)
    q, g = (torch.randn(nd_loc, H, C, device=dev, dtype=bf) for _ in range(2))
    e = torch.randn(E, H, C, device=dev, dtype=bf)
    k, v = (torch.randn(ns_loc, H, C, device=dev, dtype=bf) for _ in range(2))
    hplan = build_local_halo_plan(ei_glob, sb, db, group)
    plan = GraphCSR(hplan.local_edge_index, hplan.n_src, nd_loc)

    def step():
        qq, ee = q.detach().requires_grad_(True), e.detach().requires_grad_(True)
        kk, vv = k.detach().requires_grad_(True), v.detach().requires_grad_(True)
        ops.gt_conv_sharded(qq, kk, vv, ee, plan, hplan, group).backward(g)

    for _ in range(max(args.warmup, 3)):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    stats = torch.tensor([float(E), float(ns_loc), float(nd_loc)], device=dev, dtype=torch.float64)
    smax, ssum = stats.clone(), stats.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(smax, op=dist.ReduceOp.MAX)
    dist.all_reduce(ssum, op=dist.ReduceOp.SUM)
    ms = float(t) / args.steps
    return {"dst_split": "equal dst-row counts (tensor_split, the reference's shapes)" if weak_split == "equal" else weak_split,
            "ms_per_step": round(ms, 4), "value": float(ssum[0]) / (ms * 1e-3), "unit": UNIT, "edges_max_rank": int(smax[0]),
            "own_src_rows_max_rank": int(smax[1]), "own_src_rows_mean": int(float(ssum[1]) / world), "dst_rows_max_rank": int(smax[2])}



def _err_pair(a, b):
    """(max|a-b| / max(1, max|b|), ||a-b||_2 / ||b||_2) in fp64 on the device, chunked over rows (a, b may be 10+ GB)."""
    mx, num, den, bmax = 0.0, 0.0, 0.0, 0.0
    n = a.shape[0]
    step = max(1, (1 << 26) // max(1, a[0].numel()))
    for i in range(0, n, step):
        x, y = a[i:i + step].double(), b[i:i + step].double()
        d = x - y
        mx = max(mx, float(d.abs().max())) if d.numel() else mx
        bmax = max(bmax, float(y.abs().max())) if y.numel() else bmax
        num += float((d * d).sum())
        den += float((y * y).sum())
    return mx / max(1.0, bmax), (num ** 0.5) / max(den ** 0.5, 1e-30)


def sharded_parity(q, k_own, v_own, e, g, ei_glob, Ns_g, sb, rank, world, plan, hplan, group, dev):
    """The dst-row-sharded step (halo exchange of k / v over NVLink, gradients of halo rows sent home and added) against a
    SINGLE-RANK conv of this rank's sub-graph on the all-gathered k / v: out, dq, de compared rank by rank; dk, dv = the sum
    over ranks of the single-rank contributions (fp32 all-reduce), compared on the rows this rank owns.  Returns the maxima
    over tensors and ranks of (max-norm relative error, relative L2 error)."""
    import torch.distributed as dist

    from anemoi_models_b200 import ops
    from anemoi_models_b200.graph import GraphCSR

    H, C = q.shape[1], q.shape[2]
    ins = [t.detach().clone().requires_grad_(True) for t in (q, k_own, v_own, e)]
    out = ops.gt_conv_sharded(ins[0], ins[1], ins[2], ins[3], plan, hplan, group)
    out.backward(g)
    got = {"out": out.detach(), "dq": ins[0].grad, "dk": ins[1].grad, "dv": ins[2].grad, "de": ins[3].grad}
    # all-gather the (uneven) src shards
    sizes = [sb[r + 1] - sb[r] for r in range(world)]
    mx = max(sizes)

    def gather(t):
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([parts[r][: sizes[r]] for r in range(world)])
        del parts, pad
        return full

    k_full, v_full = gather(k_own).requires_grad_(True), gather(v_own).requires_grad_(True)
    nd_loc = q.shape[0]
    ei_ref = torch.stack([ei_glob[0], hplan.local_edge_index[1]]).contiguous()  # global src ids, local dst ids
    ref_plan = GraphCSR(ei_ref, Ns_g, nd_loc)
    qr, er = q.detach().clone().requires_grad_(True), e.detach().clone().requires_grad_(True)
    out_r = ops.gt_conv(qr, k_full, v_full, er, ref_plan)
    out_r.backward(g)
    errs = {"out": _err_pair(got["out"], out_r.detach()), "dq": _err_pair(got["dq"], qr.grad), "de": _err_pair(got["de"], er.grad)}
    del out_r, qr, er, ref_plan
    lo, hi = sb[rank], sb[rank + 1]
    for name, full in (("dk", k_full), ("dv", v_full)):
        contrib = full.grad
        mxe, num, den, bmax = 0.0, 0.0, 0.0, 0.0
        chunk = 1 << 19  # rows per fp32 all-reduce (2 GB at D = 1024)
        for r0 in range(0, Ns_g, chunk):
            r1 = min(Ns_g, r0 + chunk)
            tot = contrib[r0:r1].float()
            dist.all_reduce(tot, group=group)
            a0, a1 = max(lo, r0), min(hi, r1)
            if a1 > a0:
                x, y = got[name][a0 - lo:a1 - lo].double(), tot[a0 - r0:a1 - r0].double()
                d = x - y
                mxe, bmax = max(mxe, float(d.abs().max())), max(bmax, float(y.abs().max()))
                num += float((d * d).sum())
                den += float((y * y).sum())
            del tot
        errs[name] = (mxe / max(1.0, bmax), (num ** 0.5) / max(den ** 0.5, 1e-30))
        full.grad = None
    del k_full, v_full
    worst = torch.tensor([max(v[0] for v in errs.values()), max(v[1] for v in errs.values())], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    torch.cuda.empty_cache()
    return {"parity_checked": True, "max_rel_err": float(worst[0]), "max_rel_l2": float(worst[1]), "tolerance": 2e-2,
            "against": "single-rank conv of each rank's sub-graph on the all-gathered k / v; dk, dv summed over ranks in fp32",
            "rank0": {k: [float(f"{v[0]:.3e}"), float(f"{v[1]:.3e}")] for k, v in errs.items()}}


def config5_block(world, rank, dev, group, steps, warmup, dst_split="equal", parity=True):
    """BASELINE configs[4] / SURVEY 8e: GT mapper conv on the o1280 (6,599,680 pts) -> n320 (542,080 pts) cut-off graph, D = 1024,
    H = 16, bf16 fwd+bwd, the WHOLE graph split over `world` ranks by dst rows (strong scaling; world = 1: the whole graph on one
    GPU, ~85 GB).  dst shards: the reference's equal-count `tensor_split` (shapes.py:19-24) or edge-count balanced cut points."""
    import torch.distributed as dist

    from anemoi_models_b200 import ops
    from anemoi_models_b200 import synthetic as S
    from anemoi_models_b200.distributed.halo import aligned_src_bounds, build_local_halo_plan
    from anemoi_models_b200.distributed.shapes import tensor_split_sizes
    from anemoi_models_b200.graph import GraphCSR

    H, C = 16, 64
    Nd_g, Ns_g = 542080, S.octahedral_size(1280)
    radius = 0.6 * S.fibonacci_max_nn_distance(Nd_g)
    db = np.concatenate([[0], np.cumsum(tensor_split_sizes(Nd_g, world))]).tolist()
    if dst_split == "balanced" and world > 1:
        # in-degree of every dst: each rank counts its equal-count shard, the counts are all-gathered
        ei0, _, _, _ = S.o1280_to_n320_band(world, rank, radius=radius)
        deg = torch.from_numpy(np.bincount(ei0[1] - db[rank], minlength=db[rank + 1] - db[rank])).to(dev)
        pad = torch.zeros(max(db[r + 1] - db[r] for r in range(world)), dtype=deg.dtype, device=dev)
        pad[: deg.numel()] = deg
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        deg_all = torch.cat([parts[r][: db[r + 1] - db[r]] for r in range(world)]).cpu().numpy()
        db = S.edge_balanced_bounds(deg_all, world)
        del ei0
    ei_np, _, _, _ = S.o1280_to_n320_band(world, rank, radius=radius, dst_bounds=db)
    ei_glob = torch.from_numpy(ei_np).to(dev)
    E = ei_glob.shape[1]
    torch.manual_seed(4321 + rank)
    if world > 1:
        sb = aligned_src_bounds(ei_glob, Ns_g, group)
    else:
        sb = [0, Ns_g]
    nd_loc, ns_loc = db[rank + 1] - db[rank], sb[rank + 1] - sb[rank]
    bf = torch.bfloat16
    q = torch.randn(nd_loc, H, C, device=dev, dtype=bf)
    g = torch.randn(nd_loc, H, C, device=dev, dtype=bf)
    e = torch.randn(E, H, C, device=dev, dtype=bf)
    k_own = torch.randn(ns_loc, H, C, device=dev, dtype=bf)
    v_own = torch.randn(ns_loc, H, C, device=dev, dtype=bf)
    if world > 1:
        hplan = build_local_halo_plan(ei_glob, sb, db, group)
        plan = GraphCSR(hplan.local_edge_index, hplan.n_src, nd_loc)
    else:
        hplan, plan = None, GraphCSR(ei_glob, Ns_g, nd_loc)

    def step():
        qq, ee = q.detach().requires_grad_(True), e.detach().requires_grad_(True)
        kk, vv = k_own.detach().requires_grad_(True), v_own.detach().requires_grad_(True)
        out = ops.gt_conv_sharded(qq, kk, vv, ee, plan, hplan, group) if world > 1 else ops.gt_conv(qq, kk, vv, ee, plan)
        out.backward(g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    par = None
    if world > 1 and parity:
        par = sharded_parity(q, k_own, v_own, e, g, ei_glob, Ns_g, sb, rank, world, plan, hplan, group, dev)
    for _ in range(max(3, warmup)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    stats = torch.tensor([ev0.elapsed_time(ev1) / steps, float(E), float(hplan.n_halo if hplan is not None else 0), float(ns_loc)],
                         device=dev, dtype=torch.float64)
    mx, sm = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ab = algorithmic_bytes(float(sm[1]), Ns_g, Nd_g, 2, H * C, H)["step_survey"]
    peak, _ = peaks()
    ms = float(mx[0])
    del q, g, e, k_own, v_own, plan, hplan
    torch.cuda.empty_cache()
    return {"workload": "GT mapper conv o1280 (6,599,680) -> n320 (542,080 Fibonacci), cut-off 0.6, D=1024, H=16, bf16 fwd+bwd, whole graph "
                        f"dst-row sharded over {world} rank(s) (strong scaling), k / v halo exchange + gradient return inside the timed region",
            "dst_split": dst_split if world > 1 else "n/a", "src_split": "aligned" if world > 1 else "n/a", "scaling": "strong",
            "ms_per_step": ms, "edges_total": int(sm[1]), "edges_per_s": float(sm[1]) / (ms * 1e-3), "steps": steps,
            "edges_max_rank": int(mx[1]), "edge_imbalance_max_over_mean": float(mx[1]) / (float(sm[1]) / world),
            "halo_rows_max_rank": int(mx[2]), "own_src_rows_max_rank": int(mx[3]),
            "hbm_frac_of_peak_per_gpu": ab / world / (ms * 1e-3) / 1e9 / peak, "parity": par}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's unfused op sequence (oracle port of conv.py + PyG) on host cores
# ----------------------------------------------------------------------------------------------------------------
def reference_conv():
    """(GraphTransformerConv class, kind): the UNMODIFIED reference class from baseline/_ref (scripts/install_ref.sh; PyG through
    oracle/pyg_shim) when it travelled with the snapshot -> kind "reference"; otherwise None -> the oracle port."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "anemoi", "models")):
        return None, "port"
    for p in (os.path.join(ROOT, "oracle", "pyg_shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        from anemoi.models.layers.conv import GraphTransformerConv as RefConv
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] reference import failed ({exc!r}); timing the oracle port instead", file=sys.stderr)
        return None, "port"
    return RefConv, "reference"


def cpu_reference_sample(ei_np, Ns, Nd, steps, warmup, frac=None):
    """The reference's own CPU path, fp32, all host threads: `GraphTransformerConv.forward` + autograd backward
    (reference layers/conv.py:98-142 over PyG propagate / softmax / scatter) on the headline graph.  The unfused autograd
    graph keeps ~10 [E, D] fp32 tensors (~30 GB at the headline): with >= 64 GB of free RAM the WHOLE graph is timed
    (frac = 1, same configuration as the GPU arm), otherwise a contiguous 1/8 dst subset (src rows compacted)."""
    from oracle import gtconv as og

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if frac is None:
        try:
            import psutil

            frac = 1 if psutil.virtual_memory().available >= 64 * 2 ** 30 else 8
        except Exception:  # noqa: BLE001
            frac = 8
    if frac > 1:
        nd_s = Nd // frac
        keep = ei_np[1] < nd_s
        sub = ei_np[:, keep]
        needed, inv = np.unique(sub[0], return_inverse=True)
        ei = torch.from_numpy(np.stack([inv.astype(np.int64), sub[1]]))
        ns_s = len(needed)
    else:
        ei, ns_s, nd_s = torch.from_numpy(np.ascontiguousarray(ei_np)), Ns, Nd
    E = ei.shape[1]
    gen = torch.Generator().manual_seed(0)
    q = torch.randn(nd_s, H, C, generator=gen)
    k = torch.randn(ns_s, H, C, generator=gen)
    v = torch.randn(ns_s, H, C, generator=gen)
    e = torch.randn(E, H, C, generator=gen)
    g = torch.randn(nd_s, H, C, generator=gen)
    RefConv, kind = reference_conv()
    if RefConv is not None:
        conv = RefConv(out_channels=C)
        ins = [t.requires_grad_(True) for t in (q, k, v, e)]

        def step():
            for t in ins:
                t.grad = None
            conv(ins[0], ins[1], ins[2], ins[3], ei, size=(ns_s, nd_s)).backward(g)
        what = "anemoi.models.layers.conv.GraphTransformerConv (unmodified, baseline/_ref) forward + autograd backward over oracle/pyg_shim"
    else:
        def step():
            og.gt_conv_unfused_fwd_bwd(q, k, v, e, ei, g, (ns_s, nd_s))
        what = "oracle/gtconv.py gt_conv_unfused = reference conv.py:98-142 + PyG op sequence on torch CPU"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = ("the whole headline graph" if frac == 1 else f"contiguous 1/{frac} dst subset of the headline graph")
    return {"value": E / dt, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": round(dt * 1e3, 2), "same_config": frac == 1,
            "sample": f"{sample} ({E} edges, {ns_s} src, {nd_s} dst), fp32, {steps} timed fwd+bwd after {warmup} warm-up; {what}"}


def run_graphconv(args):
    """Report line (not the headline): one GraphConvProcessorBlock layer (reference block.py:170-223), fwd+bwd, bf16, on the
    multi-scale icosahedral mesh of BASELINE configs[2] (refinement 6: 40,962 nodes, 327,660 edges), hidden D=512.
    The edge MLP is GEMM work (tensor roofline); the kernels of this repo cover the gather/activation/LayerNorm/scatter."""
    import anemoi_models_b200 as b2
    from anemoi_models_b200 import synthetic as S

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    Dg = args.graphconv_dim
    xyz, ei_np = S.multiscale_icosahedral_mesh(6)
    N, E = xyz.shape[0], ei_np.shape[1]
    ei = torch.from_numpy(ei_np).to(dev)
    torch.manual_seed(0)
    blk = b2.GraphConvProcessorBlock(Dg, Dg).to(dev).to(torch.bfloat16)
    x = torch.randn(N, Dg, device=dev, dtype=torch.bfloat16)
    e = torch.randn(E, Dg, device=dev, dtype=torch.bfloat16)
    gx, ge = torch.randn_like(x), torch.randn_like(e)
    shapes = ([[N, Dg]], [[N, Dg]], None)

    def step():
        xx, ee = x.detach().requires_grad_(True), e.detach().requires_grad_(True)
        nodes, edges = blk(xx, ee, ei, shapes)
        torch.autograd.backward([nodes, edges], [gx, ge])

    sampler = ClockSampler(0)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    breakdown = None
    if args.profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
        total = sum(r.device_time_total for r in rows)
        breakdown = [{"kernel": r.key[:100], "ms": round(r.device_time_total / 1e3, 3), "calls": r.count,
                      "share": round(r.device_time_total / max(total, 1), 4)} for r in rows[:20]]
    # the same forward + backward captured ONCE into a CUDA graph and replayed (~100 launches of 10-300 us: the eager step is partly
    # bound by the host issuing them); every kernel of this library launches on the capturing stream, nothing synchronises
    graphed = None
    if args.cuda_graph:
        try:
            xs, es = x.detach().clone().requires_grad_(True), e.detach().clone().requires_grad_(True)
            nodes_e, edges_e = blk(xs, es, ei, shapes)  # eager result on the static inputs, for the comparison below
            nodes_e, edges_e = nodes_e.detach().clone(), edges_e.detach().clone()
            for p_ in blk.parameters():
                p_.grad = None
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                nodes_g, edges_g = blk(xs, es, ei, shapes)
                torch.autograd.backward([nodes_g, edges_g], [gx, ge])
            for _ in range(3):
                cg.replay()
            torch.cuda.synchronize()
            same = bool(torch.equal(nodes_g, nodes_e) and torch.equal(edges_g, edges_e))
            ev0.record()
            for _ in range(args.steps):
                cg.replay()
            ev1.record()
            torch.cuda.synchronize()
            graphed = {"ms_per_step": ev0.elapsed_time(ev1) / args.steps, "outputs_bit_identical_to_eager": same,
                       "what": "torch.cuda.CUDAGraph capture of the block's forward + backward, replayed"}
        except Exception as ex:  # noqa: BLE001 -- report line only
            graphed = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    # executed GEMM FLOPs with the split first layer: edge GEMMs pe, W1, W2 (3 x 2*E*D^2), node GEMMs pi, pj (2 x 2*N*D^2),
    # node MLP 2*N*(2D*D + D*D + D*D); backward = 2x forward
    flops_fwd = 6 * E * Dg * Dg + 4 * N * Dg * Dg + 8 * N * Dg * Dg
    tfs = 3 * flops_fwd / (ms * 1e-3) / 1e12
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tpeak = 1590.0  # fallback (B200_PROFILING.md) when the driver-written file is absent
    if os.path.exists(ppath):
        with open(ppath) as f:
            tpeak = float(json.load(f)["bf16_tflops"])
    line = {"metric": "graphconv_block_fwd_bwd_edges_per_s", "value": E / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"GraphConvProcessorBlock layer, multi-scale icosahedral mesh r6 (N={N}, E={E}), D={Dg}, bf16 fwd+bwd (report line)"},
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "achieved": round(tfs, 1), "peak": tpeak, "unit": "TFLOP/s", "frac": round(tfs / tpeak, 4),
                         "traffic": None, "note": "executed GEMM FLOPs (split first layer) of the whole block over the block time; GEMMs run on the tcgen05 kernel (csrc/gemm_tc.cu)"},
            "cuda_graph": graphed,
            "kernel_breakdown": breakdown}
    print(json.dumps(line), flush=True)


def run_edgepath(args):
    """Report line (not the headline; round-2 work in progress): the edge path AS THE BLOCK RUNS IT on the headline graph --
    `lin_edge(raw [E,11])` + conv forward+backward through the product path, against the same with lin_edge folded into the conv
    (ops.gt_conv_folded).  Gradients flow to q, k, v, the raw edge features and lin_edge's parameters in both."""
    from anemoi_models_b200 import ops
    from anemoi_models_b200 import synthetic as S
    from anemoi_models_b200.graph import GraphCSR

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    H, C, ed = 16, 64, 11
    which = args.edgepath_graph
    if which == "encoder":
        ei_np, ns, nd, _ = S.encoder_graph(SRC_POINTS, DST_N)
    else:
        ei_np, ns, nd, _, _ = build_shard(1, 0, which)
    ei = torch.from_numpy(ei_np).to(dev)
    E = ei.shape[1]
    plan = GraphCSR(ei, ns, nd)
    torch.manual_seed(0)
    q, k, v = (torch.randn(n, H, C, device=dev, dtype=torch.bfloat16).requires_grad_(True) for n in (nd, ns, ns))
    raw = torch.rand(E, ed, device=dev).requires_grad_(True)
    lin = torch.nn.Linear(ed, H * C).to(dev)
    g = torch.randn(nd, H, C, device=dev, dtype=torch.bfloat16)

    from anemoi_models_b200.layers.block import _lin_edge

    def product():  # what the block does by default: lin_edge on the tcgen05 GEMM (K padded to 16), then the fused conv
        e = _lin_edge(lin, raw, True)
        ops.gt_conv(q, k, v, e.view(E, H, C), plan).backward(g)

    def product_cublas():  # round-1 path: lin_edge through nn.Linear / cuBLASLt under autocast
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e = lin(raw)
        ops.gt_conv(q, k, v, e.view(E, H, C), plan).backward(g)

    def folded():
        ops.gt_conv_folded(q, k, v, raw, lin.weight, lin.bias, plan).backward(g)

    sampler = ClockSampler(0)
    res = {}
    with sampler:
        for name, fn in (("lin_edge_plus_conv", product), ("lin_edge_cublas_plus_conv", product_cublas), ("folded", folded)):
            for _ in range(max(args.warmup, 3)):
                fn()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(args.steps):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            res[name] = ev0.elapsed_time(ev1) / args.steps
    ms = res["lin_edge_plus_conv"]
    line = {"metric": "gt_block_edge_path_fwd_bwd_edges_per_s", "value": E / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"edge path of the GT block on the {which} graph (Ns={ns}, Nd={nd}): lin_edge(raw[E,11]) folded into the conv "
                                   "vs lin_edge GEMM + conv (report line)", "edges_total": int(E)},
            "clocks": sampler.summary(), "ms_product_path_lin_edge_plus_conv": res["lin_edge_plus_conv"], "ms_lin_edge_on_cublas_plus_conv": res["lin_edge_cublas_plus_conv"],
            "ms_folded": res["folded"]}
    print(json.dumps(line), flush=True)


def run_model(args, emit=True, impl="ours", recompute=None):
    """Report line (not the headline): AIFS-like n320/o96 encoder-processor-decoder training step (BASELINE configs[3]) built
    from this repo's drop-in blocks the way the reference's mappers/processor wire them (mapper.py:245-272,
    processor.py:317-343): node embeddings, trainable edge features (3 geometric + 8 trainable columns), GT mapper block
    n320->o96, 16 GT processor blocks on o96 (8-NN) in 8 checkpointed chunks, GT mapper block o96->n320 (3-NN), output
    extractor; bf16 autocast, fwd + bwd + fused AdamW step, activation checkpointing per mapper / chunk as in
    models/encoder_processor_decoder.py:159-166 and processor.py:73-77."""
    from torch.utils.checkpoint import checkpoint as _checkpoint

    import anemoi_models_b200 as b2
    from anemoi_models_b200 import synthetic as S
    from anemoi_models_b200.graph import get_csr

    # --model-recompute off: the same step WITHOUT activation checkpointing (numerically the same training step; the reference
    # checkpoints unconditionally because the model was sized for 40-80 GB GPUs -- at 180 GB the activations of this model fit)
    recompute = (args.model_recompute != "off") if recompute is None else recompute

    use_graph = bool(getattr(args, "cuda_graph", False)) and impl != "reference"

    def checkpoint(fn, *a, use_reentrant=False):
        # (no dropout anywhere: the RNG state need not be stashed -- reading it is not allowed while a CUDA graph is captured)
        return _checkpoint(fn, *a, use_reentrant=use_reentrant, preserve_rng_state=not use_graph) if recompute else fn(*a)

    if impl == "reference":
        # comparator: the SAME model wired from the UNMODIFIED reference blocks (baseline/_ref, PyG op sequence through the shim as
        # torch ops on the GPU, nn.Linear / nn.LayerNorm on cuBLASLt / ATen) -- the like-for-like "reference on the same GPU"
        ref = os.path.join(ROOT, "baseline", "_ref")
        for p_ in (os.path.join(ROOT, "oracle", "pyg_shim"), ref):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        from anemoi.models.layers import block as blocks_mod
    else:
        blocks_mod = b2

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.cuda.reset_peak_memory_stats(dev)
    torch.manual_seed(0)
    hid, heads, layers, chunks, nvar = D, H, args.model_layers, 8, 100
    hidden_xyz, _ = S.octahedral_grid(DST_N)
    data_xyz = S.fibonacci_sphere(SRC_POINTS)
    Nh, Ndata = len(hidden_xyz), SRC_POINTS
    radius = 0.6 * S.max_nn_distance(hidden_xyz)
    graphs = {"enc": (S.cutoff_edges(data_xyz, hidden_xyz, radius), Ndata, Nh),
              "proc": (S.knn_edges(hidden_xyz, hidden_xyz, 8, exclude_self=True), Nh, Nh),
              "dec": (S.knn_edges(hidden_xyz, data_xyz, 3), Nh, Ndata)}
    ei = {k_: torch.from_numpy(g_[0]).to(dev) for k_, g_ in graphs.items()}
    if impl != "reference":
        for k_, g_ in graphs.items():
            get_csr(ei[k_], g_[1], g_[2])  # one-off plan build, outside the step like the reference's constant edge buffers

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            nn = torch.nn
            self.emb_data = nn.Linear(nvar + 4, hid)
            self.emb_hidden = nn.Linear(4, hid)
            self.emb_data_dec = nn.Linear(nvar + 4, hid)
            self.edge_geo = nn.ParameterDict({k_: nn.Parameter(torch.rand(ei[k_].shape[1], 3), requires_grad=False) for k_ in ei})
            self.edge_train = nn.ParameterDict({k_: nn.Parameter(torch.randn(ei[k_].shape[1], 8) * 0.1) for k_ in ei})
            self.enc = blocks_mod.GraphTransformerMapperBlock(hid, 4 * hid, hid, edge_dim=11, num_heads=heads)
            self.proc = nn.ModuleList([blocks_mod.GraphTransformerProcessorBlock(hid, 4 * hid, hid, edge_dim=11, num_heads=heads)
                                       for _ in range(layers)])
            self.dec = blocks_mod.GraphTransformerMapperBlock(hid, 4 * hid, hid, edge_dim=11, num_heads=heads)
            self.extract = nn.Sequential(nn.LayerNorm(hid), nn.Linear(hid, nvar))

        def edge_attr(self, k_):
            return torch.cat([self.edge_geo[k_], self.edge_train[k_]], dim=-1)  # TrainableTensor (layers/graph.py:37-44)

        def forward(self, x_data, x_hidden):
            def run_enc(xd, xh):
                (s, d), _ = self.enc((self.emb_data(xd), self.emb_hidden(xh)), self.edge_attr("enc"), ei["enc"],
                                     (None, None, None), 1, size=(Ndata, Nh))
                return d

            def run_chunk(i0, x):
                ea = self.edge_attr("proc")
                for blk in self.proc[i0:i0 + layers // chunks]:
                    x, _ = blk(x, ea, ei["proc"], (None, None, None), 1)
                return x

            def run_dec(xh, xd):
                (s, d), _ = self.dec((xh, self.emb_data_dec(xd)), self.edge_attr("dec"), ei["dec"], (None, None, None), 1,
                                     size=(Nh, Ndata))
                return self.extract(d)

            latent = checkpoint(run_enc, x_data, x_hidden, use_reentrant=False)
            x = latent
            for c in range(chunks):
                x = checkpoint(run_chunk, c * (layers // chunks), x, use_reentrant=False)
            x = x + latent
            return checkpoint(run_dec, x, x_data, use_reentrant=False)

    model = Model().to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=use_graph)
    x_data = torch.randn(Ndata, nvar + 4, device=dev)
    x_hidden = torch.randn(Nh, 4, device=dev)
    target = torch.randn(Ndata, nvar, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = model(x_data, x_hidden)
        loss = (y.float() - target).square().mean()
        loss.backward()
        opt.step()
        return loss

    sampler = ClockSampler(0)
    for _ in range(max(3, min(args.warmup, 3))):
        step()
    torch.cuda.synchronize()
    n = max(1, min(args.steps, 10))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        ev0.record()
        for _ in range(n):
            loss = step()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    graphed = None
    if use_graph:  # the whole training step (forward, checkpoint recompute, backward, fused AdamW) captured once and replayed
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            loss = None
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                loss_g = step()
            for _ in range(2):
                cg.replay()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(n):
                cg.replay()
            ev1.record()
            torch.cuda.synchronize()
            graphed = {"ms_per_step": ev0.elapsed_time(ev1) / n, "loss": float(loss_g.detach()),
                       "what": "torch.cuda.CUDAGraph capture of the whole training step, replayed"}
            loss = loss_g
        except Exception as ex:  # noqa: BLE001 -- report line only
            graphed = {"error": f"{type(ex).__name__}: {ex}"[:300]}
            loss = torch.zeros(())
    breakdown = None
    if args.profile and graphed is None:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
        total = sum(r.device_time_total for r in rows)
        breakdown = [{"kernel": r.key[:90], "ms": round(r.device_time_total / 1e3, 3), "calls": r.count,
                      "share": round(r.device_time_total / max(total, 1), 4)} for r in rows[:25]]
        mine = sum(r.device_time_total for r in rows if "ab2::" in r.key)
        breakdown.insert(0, {"kernel": "ALL ab2:: kernels (this repo)", "ms": round(mine / 1e3, 3), "share": round(mine / max(total, 1), 4)})
    nparams = sum(p.numel() for p in model.parameters() if p.requires_grad)
    etot = sum(ei[k_].shape[1] * (layers if k_ == "proc" else 1) for k_ in ei)
    line = {"metric": "aifs_like_n320_o96_train_step_ms", "impl": impl, "value": ms, "unit": "ms/step", "n_gpus": 1, "steps": n, "warmup": 3,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"AIFS-like enc-proc-dec step: GT mapper n320->o96 (E={ei['enc'].shape[1]}), {layers} GT processor layers on "
                                   f"o96 8-NN (E={ei['proc'].shape[1]}) in {chunks} checkpointed chunks, GT mapper o96->n320 3-NN "
                                   f"(E={ei['dec'].shape[1]}); hidden {hid}, {heads} heads, MLP x4, bf16 autocast, fwd+bwd ("
                                   + ("with checkpoint recompute" if recompute else "NO activation checkpointing: every activation kept")
                                   + f") + fused AdamW; {nparams / 1e6:.0f} M parameters (report line)",
                       "activation_checkpointing": bool(recompute),
                       "conv_edges_per_step": int(etot), "loss": float(loss.detach()), "optimizer_steps_before_loss": 3 + n},
            "clocks": sampler.summary(), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 1),
            "cuda_graph": graphed, "kernel_breakdown": breakdown}
    if emit:
        print(json.dumps(line), flush=True)
    del model, opt
    torch.cuda.empty_cache()
    return line


def run_o1280(args):
    """Report line: BASELINE configs[4] on its own (`--workload o1280 [--dst-split balanced]`), strong scaling over --gpus."""
    import torch.distributed as dist

    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    sampler = ClockSampler(local_rank)
    with sampler:
        blk = config5_block(world, rank, dev, group, steps=args.steps, warmup=args.warmup, dst_split=args.dst_split,
                            parity=not args.no_parity_check)
    if rank == 0:
        line = {"metric": METRIC, "value": blk["edges_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": {k: blk[k] for k in ("workload", "dst_split", "src_split", "edges_total", "edges_max_rank",
                                                                     "edge_imbalance_max_over_mean", "halo_rows_max_rank", "own_src_rows_max_rank")},
                "clocks": sampler.summary(), "roofline_step": {"frac": blk["hbm_frac_of_peak_per_gpu"], "note": "SURVEY 8d step bytes / P over the step time, of the measured HBM peak"},
                "parity": blk["parity"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from anemoi_models_b200 import synthetic as S

    ei_np, Ns, Nd, _ = S.encoder_graph(SRC_POINTS, DST_N)
    res = cpu_reference_sample(ei_np, Ns, Nd, steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 5)), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(1), "note": "each step = " + res["sample"].split(",")[0] + " on host cores"},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=16, help="dst-row chunks of the streamed host-buffer call (1 = unstreamed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--src-split", default="aligned", choices=["aligned", "equal"],
                    help="multi-GPU: src row ownership aligned to the dst shards (default) or the reference's equal-count tensor_split")
    ap.add_argument("--no-model-step", action="store_true", help="encoder workload at 1 GPU: skip the AIFS-like model step block")
    ap.add_argument("--no-parity-check", action="store_true", help="multi-GPU: skip the sharded-vs-single-rank comparison made before timing")
    ap.add_argument("--config5", default="auto", choices=["auto", "on", "off"],
                    help="encoder workload: also measure BASELINE configs[4] (o1280 -> n320, whole graph over the ranks) as an extra block")
    ap.add_argument("--dst-split", default="equal", choices=["equal", "balanced"], help="o1280 workload: dst shard cut points")
    ap.add_argument("--weak-split", default="work", choices=["work", "equal"],
                    help="headline at --gpus > 1: dst cut points that give every rank the single-GPU workload's work (default), or the "
                         "reference's equal-count tensor_split (whose polar ranks have 1.13x the src rows at 8 ranks); the other "
                         "one is measured too and reported as a sub-block")
    ap.add_argument("--edgepath-graph", default="encoder", choices=["encoder", "decoder", "processor"])
    ap.add_argument("--graphconv-dim", type=int, default=512)
    ap.add_argument("--model-layers", type=int, default=16)
    ap.add_argument("--model-impl", default="ours", choices=["ours", "reference"],
                    help="model workload: this repo's blocks, or the unmodified reference blocks from baseline/_ref on the same GPU")
    ap.add_argument("--model-recompute", default="on", choices=["on", "off"],
                    help="model workload: activation checkpointing per mapper / processor chunk as the reference wires it (on), or every "
                         "activation kept in HBM (off)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="graphconv / model workloads: also capture the step into a CUDA graph and time its replay")
    ap.add_argument("--profile", action="store_true", help="model workload: add a per-kernel device-time breakdown of one step")
    ap.add_argument("--workload", default="encoder", choices=["encoder", "decoder", "processor", "graphconv", "model", "edgepath", "o1280", "config1-enc", "config1-proc", "config1-dec"],
                    help="encoder = BASELINE configs[1] (the headline); the others are extra report lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "graphconv":
        run_graphconv(args)
    elif args.workload == "model":
        run_model(args, impl=args.model_impl)
    elif args.workload == "edgepath":
        run_edgepath(args)
    elif args.workload == "o1280":
        run_o1280(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
