"""autograd wrappers over the C ABI (include/anemoi_b200.h).

Every Function is stateless per call, saves only tensors with `ctx.save_for_backward` (so non-reentrant
activation checkpointing -- reference processor.py:76, encoder_processor_decoder.py:159-166 -- works), takes the
CUDA stream of the *calling* thread (backward runs on autograd's worker thread), and fails loudly on CPU tensors:
there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .graph import GraphCSR

ACT_CODES = {"SiLU": 0, "GELU": 1, "ReLU": 2, "Identity": 3}


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("anemoi_models_b200 runs on CUDA tensors only (no CPU fallback); got a CPU tensor")


def _common_dtype(*tensors: Tensor) -> torch.dtype:
    dt = tensors[0].dtype
    for t in tensors[1:]:
        dt = torch.promote_types(dt, t.dtype)
    if dt not in (torch.float32, torch.bfloat16):
        if dt in (torch.float16,):
            return torch.float32
        raise TypeError(f"unsupported dtype {dt}: anemoi_models_b200 computes in float32 or bfloat16")
    return dt


def _aligned(t: Tensor) -> Tensor:
    """contiguous AND 16-byte aligned (the kernels use 16-byte vector / bulk loads): a contiguous view whose storage offset is
    not a multiple of 16 bytes is cloned."""
    t = t.contiguous()
    return t.clone() if (t.data_ptr() % 16) != 0 else t


def _conv_forward(q, k, v, k_halo, v_halo, e, plan):
    L = _lib.lib()
    Nd, H, C = q.shape
    n_own = k.shape[0]
    Ns = n_own + (k_halo.shape[0] if k_halo is not None else 0)
    out = torch.empty_like(q)
    lse2 = torch.empty((Nd, H), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.ab2_gtconv_fwd_halo(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(k_halo), _lib.ptr(v_halo), n_own,
                                         _lib.ptr(e), _lib.dtype_code(q.dtype), _lib.ptr(plan.rowptr), _lib.ptr(plan.col),
                                         _lib.ptr(plan.perm), Ns, Nd, plan.num_edges, H, C, _lib.ptr(out), _lib.ptr(lse2),
                                         _lib.current_stream(q.device)))
    return out, lse2


def _conv_backward(q, k, v, k_halo, v_halo, e, out, lse2, g, plan, need):
    """need = (dq, dk, dv, de) flags; returns (dq, dk, dv, de, dk_halo, dv_halo)."""
    L = _lib.lib()
    Nd, H, C = q.shape
    n_own = k.shape[0]
    Ns = n_own + (k_halo.shape[0] if k_halo is not None else 0)
    E = plan.num_edges
    g = g.contiguous()
    if g.dtype != q.dtype:
        g = g.to(q.dtype)
    dq = torch.empty_like(q) if need[0] else None
    dk = torch.empty_like(k) if need[1] else None
    dv = torch.empty_like(v) if need[2] else None
    de = torch.empty_like(e) if need[3] else None
    dkh = torch.empty_like(k_halo) if (need[1] and k_halo is not None) else None
    dvh = torch.empty_like(v_halo) if (need[2] and v_halo is not None) else None
    ws_bytes = L.ab2_gtconv_bwd_workspace_bytes(E, H)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if (need[1] or need[2]) else None
    with torch.cuda.device(q.device):
        _lib.check(L.ab2_gtconv_bwd_halo(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(k_halo), _lib.ptr(v_halo), n_own,
                                         _lib.ptr(e), _lib.dtype_code(q.dtype), _lib.ptr(plan.rowptr), _lib.ptr(plan.col),
                                         _lib.ptr(plan.perm), _lib.ptr(plan.colptr), _lib.ptr(plan.csr2csc), _lib.ptr(plan.crow),
                                         Ns, Nd, E, H, C, _lib.ptr(out), _lib.ptr(lse2), _lib.ptr(g), _lib.ptr(dq), _lib.ptr(dk),
                                         _lib.ptr(dv), _lib.ptr(dkh), _lib.ptr(dvh), _lib.ptr(de), _lib.ptr(ws),
                                         ws_bytes if ws is not None else 0, _lib.current_stream(q.device)))
    return dq, dk, dv, de, dkh, dvh


class _GTConvFn(torch.autograd.Function):
    """Fused GraphTransformerConv (reference layers/conv.py:98-142 + PyG softmax/scatter).  The src rows may come in two
    pieces (own shard + halo buffer); gradients are returned for each piece."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, e: Tensor, k_halo, v_halo, plan: GraphCSR) -> Tensor:
        out, lse2 = _conv_forward(q, k, v, k_halo, v_halo, e, plan)
        ctx.has_halo = k_halo is not None
        if ctx.has_halo:
            ctx.save_for_backward(q, k, v, e, out, lse2, k_halo, v_halo)
        else:
            ctx.save_for_backward(q, k, v, e, out, lse2)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        if ctx.has_halo:
            q, k, v, e, out, lse2, k_halo, v_halo = ctx.saved_tensors
        else:
            (q, k, v, e, out, lse2), k_halo, v_halo = ctx.saved_tensors, None, None
        n = ctx.needs_input_grad
        need = (n[0], n[1] or n[4], n[2] or n[5], n[3])
        dq, dk, dv, de, dkh, dvh = _conv_backward(q, k, v, k_halo, v_halo, e, out, lse2, g, ctx.plan, need)
        return dq, dk, dv, de, dkh, dvh, None


def _check_conv_args(query, key, value, edge_attr, plan, n_halo):
    if query.dim() != 3 or key.dim() != 3 or value.dim() != 3 or edge_attr.dim() != 3:
        raise ValueError("query/key/value/edge_attr must be [N, heads, channels]")
    if key.shape != value.shape or key.shape[1:] != query.shape[1:] or edge_attr.shape[1:] != query.shape[1:]:
        raise ValueError(f"inconsistent shapes q{tuple(query.shape)} k{tuple(key.shape)} v{tuple(value.shape)} e{tuple(edge_attr.shape)}")
    if edge_attr.shape[0] != plan.num_edges:
        raise ValueError(f"edge_attr has {edge_attr.shape[0]} rows but edge_index has {plan.num_edges} edges")
    if query.shape[0] != plan.num_dst or key.shape[0] + n_halo != plan.num_src:
        raise ValueError("node counts do not match the graph plan")


def gt_conv(query: Tensor, key: Tensor, value: Tensor, edge_attr: Tensor, plan: GraphCSR, halo=None) -> Tensor:
    """out[Nd,H,C] of the fused graph-transformer convolution; q [Nd,H,C], k/v [n_own,H,C], e [E,H,C] (original edge
    order).  `halo=(k_halo, v_halo)` appends src rows [n_own, n_own+n_halo) held in separate buffers (dst-row sharding)."""
    k_halo, v_halo = halo if halo is not None else (None, None)
    _require_cuda(query, key, value, edge_attr, k_halo, v_halo)
    _check_conv_args(query, key, value, edge_attr, plan, 0 if k_halo is None else k_halo.shape[0])
    dt = _common_dtype(query, key, value, edge_attr)
    q, k, v, e = (_aligned(t.to(dt)) for t in (query, key, value, edge_attr))
    if k_halo is not None:
        k_halo, v_halo = _aligned(k_halo.to(dt)), _aligned(v_halo.to(dt))
    out = _GTConvFn.apply(q, k, v, e, k_halo, v_halo, plan)
    # fp16 inputs are computed in fp32 (the kernels take fp32 / bf16); hand back the caller's dtype like the reference does
    return out.to(query.dtype) if query.dtype == torch.float16 else out


# ---------------------------------------------------------------------------------------------------------------------
# ROUND-2 WORK IN PROGRESS (AB2_EDGE_FOLD=1; one green GPU run, profiles/r01/fold_draft_r01ah.log): conv with lin_edge folded in -- see
# csrc/gtconv_fold.cu.  The three kernel calls are module-level functions so that the CPU tests can check this host glue
# (padding, the [Nd]-sized einsums with W, gradient assembly) against the reference with a torch emulation of the kernels.
# ---------------------------------------------------------------------------------------------------------------------
FOLD_COLS = 16  # raw edge features + one constant-1 column that carries the bias, zero padded


def _fold_fwd_kernel(q, k, v, rawp, qw, plan):
    L = _lib.lib()
    Nd, H, C = q.shape
    out = torch.empty_like(q)
    lse2 = torch.empty((Nd, H), dtype=torch.float32, device=q.device)
    R = torch.empty((Nd, H, FOLD_COLS), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.ab2_gtconv_fold_fwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(rawp), _lib.ptr(qw), _lib.dtype_code(q.dtype),
                                         _lib.ptr(plan.rowptr), _lib.ptr(plan.col), _lib.ptr(plan.perm), k.shape[0], Nd, plan.num_edges,
                                         H, C, _lib.ptr(out), _lib.ptr(lse2), _lib.ptr(R), _lib.current_stream(q.device)))
    return out, lse2, R


def _fold_bwd_kernels(q, k, v, rawp, qw, gw, out, lse2, g, plan):
    """-> dq_part [Nd,H,C], S [Nd,H,16] (already / sqrt(C)), dk, dv [Ns,H,C], draw [E,16]"""
    L = _lib.lib()
    Nd, H, C = q.shape
    Ns, E = k.shape[0], plan.num_edges
    dq = torch.empty_like(q)
    S = torch.empty((Nd, H, FOLD_COLS), dtype=torch.float32, device=q.device)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    draw = torch.empty((E, FOLD_COLS), dtype=torch.float32, device=q.device)
    ads = torch.empty((max(E, 1), H, 2), dtype=torch.float32, device=q.device)
    st = _lib.current_stream(q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.ab2_gtconv_fold_bwd_dst(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(rawp), _lib.ptr(qw), _lib.ptr(gw),
                                             _lib.dtype_code(q.dtype), _lib.ptr(plan.rowptr), _lib.ptr(plan.col), _lib.ptr(plan.perm),
                                             _lib.ptr(plan.csr2csc), Ns, Nd, E, H, C, _lib.ptr(out), _lib.ptr(lse2), _lib.ptr(g),
                                             _lib.ptr(dq), _lib.ptr(S), _lib.ptr(ads), st))
        _lib.check(L.ab2_gtconv_bwd_src(_lib.ptr(q), _lib.ptr(g), _lib.dtype_code(q.dtype), _lib.ptr(plan.colptr), _lib.ptr(plan.crow),
                                        Ns, Nd, E, H, C, _lib.ptr(ads), _lib.ptr(dk), _lib.ptr(dv), st))
        _lib.check(L.ab2_edge_raw_grad(_lib.ptr(ads), _lib.ptr(qw), _lib.ptr(gw), _lib.ptr(plan.rowptr), _lib.ptr(plan.perm),
                                       _lib.ptr(plan.csr2csc), Nd, E, H, _lib.ptr(draw), st))
    return dq, S, dk, dv, draw


def _fold_pad(raw: Tensor, weight: Tensor, bias: Optional[Tensor], H: int, C: int) -> Tuple[Tensor, Tensor]:
    """raw [E,ed] -> [E,16] fp32 with a constant-1 column at index ed; W [H*C,ed], b [H*C] -> [H,C,16] fp32 with b in column ed."""
    E, ed = raw.shape
    if ed + 1 > FOLD_COLS:
        raise ValueError(f"folded lin_edge supports at most {FOLD_COLS - 1} raw edge features, got {ed}")
    rawp = raw.new_zeros((E, FOLD_COLS), dtype=torch.float32)
    rawp[:, :ed] = raw
    rawp[:, ed] = 1.0
    Wp = weight.new_zeros((H * C, FOLD_COLS), dtype=torch.float32)
    Wp[:, :ed] = weight
    if bias is not None:
        Wp[:, ed] = bias
    return rawp, Wp.view(H, C, FOLD_COLS)


class _GTConvFoldedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, raw: Tensor, weight: Tensor, bias: Optional[Tensor], plan: GraphCSR) -> Tensor:
        Nd, H, C = q.shape
        rawp, Wp = _fold_pad(raw, weight, bias, H, C)
        qw = torch.einsum("ihc,hcm->ihm", q.float(), Wp).contiguous()
        out_part, lse2, R = _fold_fwd_kernel(q, k, v, rawp, qw, plan)
        out = (out_part.float() + torch.einsum("ihm,hcm->ihc", R, Wp)).to(q.dtype)
        ctx.save_for_backward(q, k, v, rawp, Wp, qw, out, lse2, R)
        ctx.plan, ctx.ed, ctx.has_bias = plan, raw.shape[1], bias is not None
        ctx.in_dtypes = (raw.dtype, weight.dtype, bias.dtype if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        q, k, v, rawp, Wp, qw, out, lse2, R = ctx.saved_tensors
        Nd, H, C = q.shape
        ed = ctx.ed
        g = g.contiguous().to(q.dtype)
        gw = torch.einsum("ihc,hcm->ihm", g.float(), Wp).contiguous()
        dq_part, S, dk, dv, draw = _fold_bwd_kernels(q, k, v, rawp, qw, gw, out, lse2, g, ctx.plan)
        dq = (dq_part.float() + torch.einsum("ihm,hcm->ihc", S, Wp)).to(q.dtype)
        dWp = torch.einsum("ihc,ihm->hcm", g.float(), R) + torch.einsum("ihc,ihm->hcm", q.float(), S)  # S carries 1/sqrt(C)
        dWp = dWp.reshape(H * C, FOLD_COLS)
        rd, wd, bd = ctx.in_dtypes
        dW = dWp[:, :ed].to(wd)
        db = dWp[:, ed].to(bd) if ctx.has_bias else None
        return dq, dk, dv, draw[:, :ed].to(rd), dW, db, None


def gt_conv_folded(query: Tensor, key: Tensor, value: Tensor, raw_edge_attr: Tensor, weight: Tensor, bias: Optional[Tensor],
                   plan: GraphCSR) -> Tensor:
    """`conv(q, k, v, lin_edge(raw_edge_attr))` (reference block.py:497 + conv.py:98-142) without the [E,H,C] edge tensor;
    q [Nd,H,C], k/v [Ns,H,C], raw_edge_attr [E,ed] (ed <= 15, original edge order), weight [H*C, ed], bias [H*C] or None."""
    _require_cuda(query, key, value, raw_edge_attr, weight, bias)
    if raw_edge_attr.dim() != 2 or raw_edge_attr.shape[0] != plan.num_edges:
        raise ValueError(f"raw_edge_attr must be [E={plan.num_edges}, ed], got {tuple(raw_edge_attr.shape)}")
    if weight.shape != (query.shape[1] * query.shape[2], raw_edge_attr.shape[1]):
        raise ValueError(f"weight must be [{query.shape[1] * query.shape[2]}, {raw_edge_attr.shape[1]}], got {tuple(weight.shape)}")
    if query.shape[0] != plan.num_dst or key.shape[0] != plan.num_src or key.shape != value.shape:
        raise ValueError("node counts do not match the graph plan")
    dt = _common_dtype(query, key, value)
    q, k, v = (t.to(dt).contiguous() for t in (query, key, value))
    return _GTConvFoldedFn.apply(q, k, v, raw_edge_attr, weight, bias, plan)


def _fwd_rows(q, k, v, kh, vh, e, plan, out, lse2, d0, d1, edges):
    """forward on dst rows [d0, d1) (own src rows k / v, halo rows kh / vh)."""
    if d1 <= d0:
        return
    L = _lib.lib()
    Nd, H, C = q.shape
    row, es = H * C * q.element_size(), q.element_size()
    n_own, Ns = k.shape[0], k.shape[0] + kh.shape[0]
    _lib.check(L.ab2_gtconv_fwd_halo(q.data_ptr() + d0 * row, _lib.ptr(k), _lib.ptr(v), _lib.ptr(kh), _lib.ptr(vh), n_own, _lib.ptr(e),
                                     _lib.dtype_code(q.dtype), plan.rowptr.data_ptr() + d0 * 4, _lib.ptr(plan.col), _lib.ptr(plan.perm),
                                     Ns, d1 - d0, max(int(edges), 1), H, C, out.data_ptr() + d0 * row, lse2.data_ptr() + d0 * H * 4,
                                     _lib.current_stream(q.device)))


def _bwd_dst_rows(q, k, v, kh, vh, e, plan, out, lse2, g, dq, de, ws, d0, d1, edges):
    if d1 <= d0:
        return
    L = _lib.lib()
    Nd, H, C = q.shape
    row = H * C * q.element_size()
    n_own, Ns = k.shape[0], k.shape[0] + kh.shape[0]
    _lib.check(L.ab2_gtconv_bwd_dst_halo(q.data_ptr() + d0 * row, _lib.ptr(k), _lib.ptr(v), _lib.ptr(kh), _lib.ptr(vh), n_own,
                                         _lib.ptr(e), _lib.dtype_code(q.dtype), plan.rowptr.data_ptr() + d0 * 4, _lib.ptr(plan.col),
                                         _lib.ptr(plan.perm), _lib.ptr(plan.csr2csc), Ns, d1 - d0, max(int(edges), 1), H, C,
                                         out.data_ptr() + d0 * row, lse2.data_ptr() + d0 * H * 4, g.data_ptr() + d0 * row,
                                         dq.data_ptr() + d0 * row, _lib.ptr(de), _lib.ptr(ws), ws.numel(),
                                         _lib.current_stream(q.device)))


def _bwd_src_rows(q, g, plan, n_own, Ns, ws, dk, dv, dkh, dvh, r0, r1):
    if r1 <= r0:
        return
    L = _lib.lib()
    Nd, H, C = q.shape
    _lib.check(L.ab2_gtconv_bwd_src_range_halo(_lib.ptr(q), _lib.ptr(g), _lib.dtype_code(q.dtype), _lib.ptr(plan.colptr),
                                               _lib.ptr(plan.crow), n_own, Ns, Nd, plan.num_edges, H, C, _lib.ptr(ws), _lib.ptr(dk),
                                               _lib.ptr(dv), _lib.ptr(dkh), _lib.ptr(dvh), r0, r1, _lib.current_stream(q.device)))


import os as _os

_TRACE = _os.environ.get("AB2_TRACE", "0") == "1"
_trace_events = []  # [(label, event)] of the last overlapped forward+backward (debug: AB2_TRACE=1)


def _mark(label, device):
    if _TRACE:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        _trace_events.append((label, ev))


def trace_summary():
    """ms between consecutive marks of the last step (call after torch.cuda.synchronize())."""
    out = [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(_trace_events[:-1], _trace_events[1:])]
    _trace_events.clear()
    return out


# AB2_BOUNDARY_STREAM=0: the boundary rows run on the main stream after / before the interior rows (round-2 first version).
# Default: they run on the exchange's side stream, next to the interior kernels -- the boundary launches are small (0.02-0.075 ms
# each at 8 GPUs, a few hundred dst rows: they cannot fill the GPU on their own) and sat on the critical path of every rank.
_BOUNDARY_STREAM = _os.environ.get("AB2_BOUNDARY_STREAM", "1") != "0"


class _GTConvShardedOverlapFn(torch.autograd.Function):
    """dst-row-sharded conv with the NVLink peer-memory exchange hidden behind the interior rows.
    forward : side stream pushes k / v halo rows and, once the peers' rows have landed, runs the forward on the boundary dst rows |
              main stream runs the forward on the interior dst rows;
    backward: side stream runs the dst pass on the boundary rows and the src pass on the halo rows and sends their gradients |
              main stream runs the interior dst pass and the src pass of the own rows; then the peers' contributions are added."""

    @staticmethod
    def forward(ctx, q, k, v, e, plan: GraphCSR, hplan, group, px):
        Nd, H, C = q.shape
        out = torch.empty_like(q)
        lse2 = torch.empty((Nd, H), dtype=torch.float32, device=q.device)
        rng = px.ranges(plan)
        with torch.cuda.device(q.device):
            _mark("fwd:start", q.device)
            k_halo, v_halo, landed = px.forward_async(k, v)
            _fwd_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, *rng["interior"])
            _mark("fwd:interior", q.device)
            if _BOUNDARY_STREAM:
                with torch.cuda.stream(px.stream):  # behind the push and the wait for the peers' rows, in stream order
                    for blk in rng["boundary"]:
                        _fwd_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, *blk)
                    bdone = torch.cuda.Event()
                    bdone.record(px.stream)
                torch.cuda.current_stream(q.device).wait_event(bdone)
                _mark("fwd:wait_boundary", q.device)
            else:
                torch.cuda.current_stream(q.device).wait_event(landed)
                _mark("fwd:wait_halo", q.device)
                for blk in rng["boundary"]:
                    _fwd_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, *blk)
                _mark("fwd:boundary", q.device)
        ctx.save_for_backward(q, k, v, e, out, lse2, k_halo, v_halo)
        ctx.plan, ctx.hplan, ctx.group, ctx.px = plan, hplan, group, px
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, v, e, out, lse2, k_halo, v_halo = ctx.saved_tensors
        plan, px = ctx.plan, ctx.px
        L = _lib.lib()
        Nd, H, C = q.shape
        n_own, Ns = k.shape[0], k.shape[0] + k_halo.shape[0]
        g = g.contiguous()
        if g.dtype != q.dtype:
            g = g.to(q.dtype)
        dq, dk, dv, de = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(e)
        dkh, dvh = torch.empty_like(k_halo), torch.empty_like(v_halo)
        ws = torch.empty(L.ab2_gtconv_bwd_workspace_bytes(plan.num_edges, H), dtype=torch.uint8, device=q.device)
        rng = px.ranges(plan)
        with torch.cuda.device(q.device):
            _mark("bwd:start", q.device)
            main = torch.cuda.current_stream(q.device)
            if _BOUNDARY_STREAM:
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(px.stream):
                    px.stream.wait_event(ready)
                    for blk in rng["boundary"]:
                        _bwd_dst_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, g, dq, de, ws, *blk)
                    bdst = torch.cuda.Event()  # the own-row src pass reads the softmax-weight workspace of the boundary edges too
                    bdst.record(px.stream)
                    _bwd_src_rows(q, g, plan, n_own, Ns, ws, dk, dv, dkh, dvh, n_own, Ns)  # halo rows: only boundary edges touch them
                    parity, pushed = px.backward_async(dkh, dvh)
                _bwd_dst_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, g, dq, de, ws, *rng["interior"])
                _mark("bwd:interior_dst", q.device)
                main.wait_event(bdst)
            else:
                for blk in rng["boundary"]:
                    _bwd_dst_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, g, dq, de, ws, *blk)
                _bwd_src_rows(q, g, plan, n_own, Ns, ws, dk, dv, dkh, dvh, n_own, Ns)  # halo rows: only boundary edges touch them
                _mark("bwd:boundary+halo_src", q.device)
                parity, pushed = px.backward_async(dkh, dvh)
                _bwd_dst_rows(q, k, v, k_halo, v_halo, e, plan, out, lse2, g, dq, de, ws, *rng["interior"])
                _mark("bwd:interior_dst", q.device)
            _bwd_src_rows(q, g, plan, n_own, Ns, ws, dk, dv, dkh, dvh, 0, n_own)
            _mark("bwd:own_src", q.device)
            torch.cuda.current_stream(q.device).wait_event(pushed)
            _mark("bwd:wait_push", q.device)
            px.backward_finish(parity, dk, dv)
            _mark("bwd:rows_add", q.device)
        return dq, dk, dv, de, None, None, None, None


class _GTConvShardedFn(torch.autograd.Function):
    """dst-row-sharded conv with the halo exchange inside: forward pulls the k / v rows of peers this rank's edges
    reference (NCCL all-to-all over NVLink) straight into the halo buffers the kernels read; backward sends the
    gradients of those rows home and adds them into dk / dv in place (no full-size temporaries)."""

    @staticmethod
    def forward(ctx, q, k, v, e, plan: GraphCSR, hplan, group):
        from .distributed.halo import exchange_rows
        from .distributed.peer import get_peer_exchange

        px = get_peer_exchange(hplan, group, k.shape[1] * k.shape[2] * k.element_size(), k.device)
        if px is not None:  # NVLink peer-memory push (one kernel + a stream-ordered barrier)
            k_halo, v_halo = px.forward(k, v)
        else:  # NCCL all-to-all
            k_halo = exchange_rows(k, hplan, group)
            v_halo = exchange_rows(v, hplan, group)
        out, lse2 = _conv_forward(q, k, v, k_halo, v_halo, e, plan)
        ctx.save_for_backward(q, k, v, e, out, lse2, k_halo, v_halo)
        ctx.plan, ctx.hplan, ctx.group, ctx.px = plan, hplan, group, px
        return out

    @staticmethod
    def backward(ctx, g):
        from .distributed.halo import return_rows

        q, k, v, e, out, lse2, k_halo, v_halo = ctx.saved_tensors
        n = ctx.needs_input_grad
        dq, dk, dv, de, dkh, dvh = _conv_backward(q, k, v, k_halo, v_halo, e, out, lse2, g, ctx.plan, (n[0], n[1], n[2], n[3]))
        if ctx.px is not None and n[1] and n[2]:
            ctx.px.backward(dkh, dvh, dk, dv)
        else:
            if n[1]:
                return_rows(dkh, ctx.hplan, ctx.group, dk)
            if n[2]:
                return_rows(dvh, ctx.hplan, ctx.group, dv)
        return dq, dk, dv, de, None, None, None


def gt_conv_sharded(query: Tensor, key: Tensor, value: Tensor, edge_attr: Tensor, plan: GraphCSR, hplan, group) -> Tensor:
    """Fused conv on this rank's dst rows; key/value are the rank's own src rows, the halo is exchanged inside."""
    _require_cuda(query, key, value, edge_attr)
    _check_conv_args(query, key, value, edge_attr, plan, hplan.n_halo)
    dt = _common_dtype(query, key, value, edge_attr)
    q, k, v, e = (t.to(dt).contiguous() for t in (query, key, value, edge_attr))
    if all(t.requires_grad or not torch.is_grad_enabled() for t in (q, k, v, e)) and q.shape[0] > 0:
        from .distributed.peer import get_peer_exchange

        px = get_peer_exchange(hplan, group, k.shape[1] * k.shape[2] * k.element_size(), k.device)
        if px is not None and px.overlap:
            return _GTConvShardedOverlapFn.apply(q, k, v, e, plan, hplan, group, px)
    return _GTConvShardedFn.apply(q, k, v, e, plan, hplan, group)


class _EdgeGatherAddActFn(torch.autograd.Function):
    """h0[t] = act(pi[dst_t] + pj[src_t] + pe[t]) -- the first edge-MLP layer of GraphConv after splitting
    Linear(3D->D) over cat[x_i, x_j, e] (reference conv.py:69, mlp.py:74) into node-side and edge-side terms."""

    @staticmethod
    def forward(ctx, pi: Tensor, pj: Tensor, pe: Tensor, plan: GraphCSR, act: int) -> Tensor:
        L = _lib.lib()
        E, D = pe.shape
        dt = _lib.dtype_code(pe.dtype)
        h0 = torch.empty_like(pe)
        pre = torch.empty_like(pe)
        with torch.cuda.device(pe.device):
            _lib.check(L.ab2_edge_gather_add_act(_lib.ptr(pi), _lib.ptr(pj), _lib.ptr(pe), _lib.ptr(plan.edge_index), E,
                                                 plan.num_src, plan.num_dst, D, dt, act, _lib.ptr(h0), _lib.ptr(pre),
                                                 _lib.current_stream(pe.device)))
        ctx.save_for_backward(pre)
        ctx.plan, ctx.act = plan, act
        return h0

    @staticmethod
    def backward(ctx, g: Tensor):
        (pre,) = ctx.saved_tensors
        plan: GraphCSR = ctx.plan
        L = _lib.lib()
        E, D = pre.shape
        dt = _lib.dtype_code(pre.dtype)
        g = g.contiguous().to(pre.dtype)
        gpe = torch.empty_like(pre)
        dpi = torch.empty((plan.num_dst, D), dtype=pre.dtype, device=pre.device)
        dpj = torch.empty((plan.num_src, D), dtype=pre.dtype, device=pre.device)
        with torch.cuda.device(pre.device):
            _lib.check(L.ab2_edge_gather_add_act_bwd(_lib.ptr(g), _lib.ptr(pre), _lib.ptr(plan.rowptr), _lib.ptr(plan.perm),
                                                     _lib.ptr(plan.colptr), _lib.ptr(plan.cpos), E, plan.num_src, plan.num_dst,
                                                     D, dt, ctx.act, _lib.ptr(gpe), _lib.ptr(dpi), _lib.ptr(dpj),
                                                     _lib.current_stream(pre.device)))
        return dpi, dpj, gpe, None, None


def edge_gather_add_act(pi: Tensor, pj: Tensor, pe: Tensor, plan: GraphCSR, activation: str) -> Tensor:
    _require_cuda(pi, pj, pe)
    dt = _common_dtype(pi, pj, pe)
    return _EdgeGatherAddActFn.apply(pi.to(dt).contiguous(), pj.to(dt).contiguous(), pe.to(dt).contiguous(), plan,
                                     ACT_CODES[activation])


class _EdgeLnResSegsumFn(torch.autograd.Function):
    """edges_new = LayerNorm(y) + e ; out = segment-sum of edges_new over dst  (reference mlp.py:84 AutocastLayerNorm,
    conv.py:69 residual, conv.py:74 scatter-sum) in one pass over the CSR."""

    @staticmethod
    def forward(ctx, y: Tensor, e: Tensor, gamma: Tensor, beta: Tensor, eps: float, plan: GraphCSR):
        L = _lib.lib()
        E, D = y.shape
        dt = _lib.dtype_code(y.dtype)
        edges_new = torch.empty_like(y)
        out = torch.empty((plan.num_dst, D), dtype=y.dtype, device=y.device)
        mean = torch.empty(E, dtype=torch.float32, device=y.device)
        rstd = torch.empty(E, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(L.ab2_edge_ln_res_segsum(_lib.ptr(y), _lib.ptr(e), _lib.ptr(gamma), _lib.ptr(beta), eps,
                                                _lib.ptr(plan.rowptr), _lib.ptr(plan.perm), E, plan.num_dst, D, dt,
                                                _lib.ptr(edges_new), _lib.ptr(out), _lib.ptr(mean), _lib.ptr(rstd),
                                                _lib.current_stream(y.device)))
        ctx.save_for_backward(y, gamma, mean, rstd)
        ctx.plan = plan
        ctx.param_dtypes = (gamma.dtype, beta.dtype)
        return edges_new, out

    @staticmethod
    def backward(ctx, g_edges: Optional[Tensor], g_out: Optional[Tensor]):
        y, gamma, mean, rstd = ctx.saved_tensors
        plan: GraphCSR = ctx.plan
        L = _lib.lib()
        E, D = y.shape
        dt = _lib.dtype_code(y.dtype)
        if g_out is None:
            g_out = torch.zeros((plan.num_dst, D), dtype=y.dtype, device=y.device)
        g_out = g_out.contiguous().to(y.dtype)
        if g_edges is not None:
            g_edges = g_edges.contiguous().to(y.dtype)
        dy = torch.empty_like(y)
        de = torch.empty_like(y)
        nparts = L.ab2_ln_bwd_parts()
        partial = torch.empty((nparts, 2, D), dtype=torch.float32, device=y.device)
        dgamma = torch.empty(D, dtype=torch.float32, device=y.device)
        dbeta = torch.empty(D, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(L.ab2_edge_ln_res_segsum_bwd(_lib.ptr(g_edges), _lib.ptr(g_out), _lib.ptr(y), _lib.ptr(gamma),
                                                    _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(plan.edge_index), E, plan.num_dst,
                                                    D, dt, _lib.ptr(dy), _lib.ptr(de), _lib.ptr(partial), nparts,
                                                    _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.current_stream(y.device)))
        return dy, de, dgamma.to(ctx.param_dtypes[0]), dbeta.to(ctx.param_dtypes[1]), None, None


def edge_ln_res_segsum(y: Tensor, e: Tensor, gamma: Tensor, beta: Tensor, eps: float, plan: GraphCSR) -> Tuple[Tensor, Tensor]:
    _require_cuda(y, e, gamma, beta)
    dt = _common_dtype(y, e)
    return _EdgeLnResSegsumFn.apply(y.to(dt).contiguous(), e.to(dt).contiguous(), gamma.to(dt).contiguous(),
                                    beta.to(dt).contiguous(), float(eps), plan)


def host_stream_meta(plan: GraphCSR, nchunks: int = 16) -> Tensor:
    """Chunk table for `ab2_gtconv_fwd_bwd_host_streamed` (host int64 [nchunks, 8], see include/anemoi_b200.h).
    One-off per plan (cached): chunk bounds by dst rows, the edge range of each chunk, the running maximum of the src ids
    referenced so far and the first src id that still has an edge in a later chunk."""
    key = int(nchunks)
    if key in plan._host_meta:
        return plan._host_meta[key]
    from .distributed.shapes import tensor_split_sizes

    Nd, Ns = plan.num_dst, plan.num_src
    nchunks = max(1, min(int(nchunks), max(Nd, 1)))
    bounds = [0]
    for s_ in tensor_split_sizes(Nd, nchunks):
        bounds.append(bounds[-1] + s_)
    rp = plan.rowptr[torch.tensor(bounds, device=plan.device)].tolist()
    smin, smax = [], []
    for c in range(nchunks):
        if rp[c + 1] > rp[c]:
            seg = plan.col[rp[c]:rp[c + 1]]
            smin.append(int(seg.min()))
            smax.append(int(seg.max()))
        else:
            smin.append(Ns)
            smax.append(-1)
    meta = torch.zeros((nchunks, 8), dtype=torch.int64)
    run_max, suffix_min = -1, [Ns] * (nchunks + 1)
    for c in range(nchunks - 1, -1, -1):
        suffix_min[c] = min(suffix_min[c + 1], smin[c])
    for c in range(nchunks):
        run_max = max(run_max, smax[c])
        meta[c, 0], meta[c, 1], meta[c, 2], meta[c, 3] = bounds[c], bounds[c + 1], rp[c], rp[c + 1]
        meta[c, 4], meta[c, 5] = run_max, suffix_min[c + 1]
    plan._host_meta[key] = meta
    return meta


def gt_conv_host(q: Tensor, k: Tensor, v: Tensor, e: Tensor, g: Tensor, plan: GraphCSR, dev_ws: Optional[Tensor] = None,
                 outs=None, nchunks: int = 16):
    """GraphTransformerConv forward+backward on PINNED HOST tensors through the host-buffer C-ABI calls.
    Dst-sorted edge lists take the streamed entry point (uploads, kernels and downloads overlap); any other edge order
    takes `ab2_gtconv_fwd_bwd_host`.  Returns pinned host tensors (out, dq, dk, dv, de)."""
    for t in (q, k, v, e, g):
        if t.is_cuda or not t.is_pinned():
            raise ValueError("gt_conv_host expects pinned host tensors")
    L = _lib.lib()
    Nd, H, C = q.shape
    Ns, E = k.shape[0], e.shape[0]
    dt = _lib.dtype_code(q.dtype)
    need = L.ab2_gtconv_host_workspace_bytes(Ns, Nd, E, H, C, dt)
    if dev_ws is None or dev_ws.numel() < need:
        dev_ws = torch.empty(need, dtype=torch.uint8, device=plan.device)
    if outs is None:
        outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (q, q, k, v, e)]
    common = [_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(e), _lib.ptr(g), dt, _lib.ptr(plan.rowptr), _lib.ptr(plan.col),
              _lib.ptr(plan.perm), _lib.ptr(plan.colptr), _lib.ptr(plan.csr2csc), _lib.ptr(plan.crow), Ns, Nd, E, H, C,
              *[_lib.ptr(o) for o in outs]]
    with torch.cuda.device(plan.device):
        if plan.perm_is_identity and nchunks > 1 and E > 0:
            meta = host_stream_meta(plan, nchunks)
            _lib.check(L.ab2_gtconv_fwd_bwd_host_streamed(*common, meta.data_ptr(), meta.shape[0], _lib.ptr(dev_ws),
                                                          dev_ws.numel(), _lib.current_stream(plan.device)))
        else:
            _lib.check(L.ab2_gtconv_fwd_bwd_host(*common, _lib.ptr(dev_ws), dev_ws.numel(), _lib.current_stream(plan.device)))
    return tuple(outs)
