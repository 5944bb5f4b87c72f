"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's edge path.

Each function cites the reference lines it follows (paths relative to
/root/reference/src/anemoi/models/).  The PyG pieces (propagate / softmax / scatter) are restated from
the published torch-geometric 2.3/2.4 API, see oracle/pyg_shim/.

Three independent formulations are kept on purpose:
  * `gt_conv_unfused`   -- the reference's op sequence, op by op, on torch CPU tensors (autograd gives
                           the backward exactly as the reference gets it).  This is also the "port" that
                           bench.py times as the CPU baseline.
  * `gt_conv_csr_f64`   -- the fused CSR-ordered forward/backward formulas the CUDA kernels implement,
                           in numpy float64 (explicit gradient formulas, no autograd).
  * `gt_conv_loops_f64` -- a per-destination pure-Python formulation for small cases (dense softmax per
                           node), used to cross-check the two above.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor


# --------------------------------------------------------------------------------------------------
# integer side: sizes, CSR
# --------------------------------------------------------------------------------------------------
def infer_size(query: Tensor, key: Tensor, size) -> Tuple[int, int]:
    """PyG `_set_size` (called from propagate, conv.py:110): size=(Ns, Nd); None entries are filled from
    the tensors, a mismatch raises ValueError."""
    ns, nd = (None, None) if size is None else (size[0], size[1])
    if ns is None:
        ns = key.shape[0]
    elif ns != key.shape[0]:
        raise ValueError(f"Encountered tensor with size {key.shape[0]} in dimension 0, but expected size {ns}.")
    if nd is None:
        nd = query.shape[0]
    elif nd != query.shape[0]:
        raise ValueError(f"Encountered tensor with size {query.shape[0]} in dimension 0, but expected size {nd}.")
    return int(ns), int(nd)


def csr_build(edge_index, num_dst: int):
    """dst-sorted CSR of `edge_index [2,E]` (row 0 = src j, row 1 = dst i; conv.py:98-121 / PyG flow
    source_to_target).  Stable: edges of one dst keep their original relative order, i.e. exactly
    `torch.sort(edge_index[1], stable=True)`.

    Returns (rowptr[int32, Nd+1], col[int32, E] = src of the p-th sorted edge, perm[int32, E] = original edge id)."""
    ei = np.asarray(edge_index.cpu().numpy() if isinstance(edge_index, Tensor) else edge_index)
    dst = ei[1].astype(np.int64)
    perm = np.argsort(dst, kind="stable")
    counts = np.bincount(dst, minlength=num_dst)
    rowptr = np.zeros(num_dst + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = ei[0][perm]
    return rowptr.astype(np.int32), col.astype(np.int32), perm.astype(np.int32)


def csc_of_csr(rowptr: np.ndarray, col: np.ndarray, num_src: int):
    """src-sorted view of the CSR positions (used by the src-side backward pass): stable sort of CSR
    positions p by col[p].  Returns (colptr[int32, Ns+1], pos[int32, E] = CSR position, row[int32, E] = dst)."""
    E = col.shape[0]
    nd = rowptr.shape[0] - 1
    row_of_p = np.repeat(np.arange(nd, dtype=np.int64), np.diff(rowptr.astype(np.int64)))
    pos = np.argsort(col.astype(np.int64), kind="stable")
    counts = np.bincount(col.astype(np.int64), minlength=num_src)
    colptr = np.zeros(num_src + 1, dtype=np.int64)
    np.cumsum(counts, out=colptr[1:])
    assert pos.shape[0] == E
    return colptr.astype(np.int32), pos.astype(np.int32), row_of_p[pos].astype(np.int32)


# --------------------------------------------------------------------------------------------------
# GraphTransformerConv -- reference op sequence (conv.py:98-142 + PyG propagate/softmax/scatter)
# --------------------------------------------------------------------------------------------------
def segment_softmax(alpha: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """PyG utils.softmax(src, index, ptr=None, num_nodes, dim=0): max on src.detach() via scatter amax
    (empty segments stay 0, include_self=False), exp(src - max[index]), divide by (scatter_sum + 1e-16)[index]."""
    idx = index.view(-1, *([1] * (alpha.dim() - 1))).expand_as(alpha)
    size = (num_nodes,) + tuple(alpha.shape[1:])
    m = alpha.new_zeros(size).scatter_reduce_(0, idx, alpha.detach(), reduce="amax", include_self=False)
    o = (alpha - m.index_select(0, index)).exp()
    s = alpha.new_zeros(size).scatter_add_(0, idx, o) + 1e-16
    return o / s.index_select(0, index)


def gt_conv_unfused(
    query: Tensor, key: Tensor, value: Tensor, edge_attr: Tensor, edge_index: Tensor, size=None,
    out_channels: Optional[int] = None,
) -> Tensor:
    """GraphTransformerConv.forward (conv.py:98-121) -> propagate -> message (conv.py:123-142) ->
    default `add` aggregation, op for op.  `out_channels` is the per-head width C used in the 1/sqrt(C)
    scale (conv.py:137 uses self.out_channels, which block.py:339 sets to C)."""
    ns, nd = infer_size(query, key, size)
    heads = query.shape[1]
    C = query.shape[2] if out_channels is None else out_channels
    src, dst = edge_index[0], edge_index[1]
    query_i = query.index_select(0, dst)  # PyG _lift
    key_j = key.index_select(0, src)
    value_j = value.index_select(0, src)
    key_j = key_j + edge_attr  # conv.py:135
    alpha = (query_i * key_j).sum(dim=-1) / C**0.5  # conv.py:137
    alpha = segment_softmax(alpha, dst, nd)  # conv.py:139
    msg = (value_j + edge_attr) * alpha.view(-1, heads, 1)  # conv.py:142 (dropout p=0 is the identity)
    idx = dst.view(-1, 1, 1).expand_as(msg)
    return msg.new_zeros((nd,) + tuple(msg.shape[1:])).scatter_add_(0, idx, msg)  # aggr="add" (conv.py:92)


def gt_conv_unfused_fwd_bwd(q, k, v, e, edge_index, g, size=None):
    """Forward + autograd backward of `gt_conv_unfused`; returns dict(out, dq, dk, dv, de)."""
    q, k, v, e = (t.detach().clone().requires_grad_(True) for t in (q, k, v, e))
    out = gt_conv_unfused(q, k, v, e, edge_index, size)
    out.backward(g)
    return {"out": out.detach(), "dq": q.grad, "dk": k.grad, "dv": v.grad, "de": e.grad}


# --------------------------------------------------------------------------------------------------
# GraphTransformerConv -- the fused CSR formulas (what the CUDA kernels compute), numpy float64
# --------------------------------------------------------------------------------------------------
def gt_conv_csr_f64(q, k, v, e, edge_index, g=None, num_dst: Optional[int] = None):
    """Per dst i, head h, incoming edge t=(j->i):  kk=k_j+e_t, vv=v_j+e_t, s_t=q_i.kk/sqrt(C),
    m=max_t s_t, l=sum_t exp(s_t-m), a_t=exp(s_t-m)/(l+1e-16), out_i=sum_t a_t vv_t, lse=m+log(l+1e-16).
    Backward (g = d out):  Dl_i=g_i.out_i, gv_t=g_i.vv_t, ds_t=a_t(gv_t-Dl_i);
    dq_i=sum_t ds_t kk_t/sqrt(C);  de_t=a_t g_i + ds_t q_i/sqrt(C);  dk_j+=ds_t q_i/sqrt(C);  dv_j+=a_t g_i.
    (No gradient flows through the max: PyG takes it on src.detach().)"""
    f = lambda t: np.asarray(t.detach().cpu().to(torch.float64).numpy() if isinstance(t, Tensor) else t, dtype=np.float64)
    q, k, v, e = f(q), f(k), f(v), f(e)
    ei = np.asarray(edge_index.cpu().numpy() if isinstance(edge_index, Tensor) else edge_index).astype(np.int64)
    nd = q.shape[0] if num_dst is None else num_dst
    H, C = q.shape[1], q.shape[2]
    src, dst = ei[0], ei[1]
    scale = 1.0 / math.sqrt(C)
    kk = k[src] + e
    vv = v[src] + e
    s = (q[dst] * kk).sum(-1) * scale  # [E,H]
    m = np.full((nd, H), -np.inf)
    np.maximum.at(m, dst, s)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    p = np.exp(s - m_safe[dst])
    l = np.zeros((nd, H))
    np.add.at(l, dst, p)
    a = p / (l[dst] + 1e-16)
    out = np.zeros((nd, H, C))
    np.add.at(out, dst, a[..., None] * vv)
    lse = m_safe + np.log(l + 1e-16)
    res = {"out": out, "lse": lse, "alpha": a}
    if g is not None:
        g = f(g)
        Dl = (g * out).sum(-1)  # [Nd,H]
        gv = (g[dst] * vv).sum(-1)  # [E,H]
        ds = a * (gv - Dl[dst])
        dq = np.zeros_like(q)
        np.add.at(dq, dst, ds[..., None] * kk * scale)
        de = a[..., None] * g[dst] + ds[..., None] * q[dst] * scale
        dk = np.zeros_like(k)
        np.add.at(dk, src, ds[..., None] * q[dst] * scale)
        dv = np.zeros_like(v)
        np.add.at(dv, src, a[..., None] * g[dst])
        res.update(dq=dq, dk=dk, dv=dv, de=de, ds=ds)
    return res


def gt_conv_loops_f64(q, k, v, e, edge_index, num_dst: Optional[int] = None) -> np.ndarray:
    """Independent small-case formulation: for every dst node build the list of incoming edges and run a
    dense softmax over it (pure-Python loops; sizes of a few hundred edges only)."""
    f = lambda t: np.asarray(t.detach().cpu().to(torch.float64).numpy() if isinstance(t, Tensor) else t, dtype=np.float64)
    q, k, v, e = f(q), f(k), f(v), f(e)
    ei = np.asarray(edge_index.cpu().numpy() if isinstance(edge_index, Tensor) else edge_index)
    nd = q.shape[0] if num_dst is None else num_dst
    H, C = q.shape[1], q.shape[2]
    out = np.zeros((nd, H, C))
    incoming = [[] for _ in range(nd)]
    for t in range(ei.shape[1]):
        incoming[int(ei[1, t])].append(t)
    for i in range(nd):
        if not incoming[i]:
            continue
        for h in range(H):
            logits = np.array([np.dot(q[i, h], k[int(ei[0, t]), h] + e[t, h]) / math.sqrt(C) for t in incoming[i]])
            w = np.exp(logits - logits.max())
            w = w / (w.sum() + 1e-16)
            for wt, t in zip(w, incoming[i]):
                out[i, h] += wt * (v[int(ei[0, t]), h] + e[t, h])
    return out


# --------------------------------------------------------------------------------------------------
# GraphConv -- reference op sequence (conv.py:27-76, mlp.py:74-84, utils.py:27-39)
# --------------------------------------------------------------------------------------------------
_ACT = {"SiLU": torch.nn.functional.silu, "GELU": torch.nn.functional.gelu, "ReLU": torch.relu, "Tanh": torch.tanh}


def mlp_forward(x: Tensor, p: dict, prefix: str, n_extra_layers: int = 0, activation: str = "SiLU") -> Tensor:
    """MLP (mlp.py:74-84): Linear,act,(Linear,act)x(n_extra+1),Linear,AutocastLayerNorm.  `p` holds the
    state_dict entries `<prefix>model.<i>.{weight,bias}`."""
    act = _ACT[activation]
    F = torch.nn.functional
    idx = 0
    h = act(F.linear(x, p[f"{prefix}model.{idx}.weight"], p[f"{prefix}model.{idx}.bias"]))
    idx += 2
    for _ in range(n_extra_layers + 1):
        h = act(F.linear(h, p[f"{prefix}model.{idx}.weight"], p[f"{prefix}model.{idx}.bias"]))
        idx += 2
    h = F.linear(h, p[f"{prefix}model.{idx}.weight"], p[f"{prefix}model.{idx}.bias"])
    idx += 1
    w, b = p[f"{prefix}model.{idx}.weight"], p[f"{prefix}model.{idx}.bias"]
    return F.layer_norm(h, (h.shape[-1],), w, b, 1e-5).type_as(h)  # AutocastLayerNorm (utils.py:33-39)


def graph_conv_unfused(x, edge_attr: Tensor, edge_index: Tensor, p: dict, prefix: str = "edge_mlp.",
                       n_extra_layers: int = 0, activation: str = "SiLU", size=None):
    """GraphConv.forward (conv.py:61-66): x_i = x_dst[dst], x_j = x_src[src];
    edges_new = edge_mlp(cat[x_i, x_j, e], 1) + e (conv.py:69); out = scatter_sum(edges_new, dst, dim_size) (conv.py:74).
    PyG overrides dim_size with size_i = Nd."""
    x_src, x_dst = (x, x) if isinstance(x, Tensor) else x
    ns, nd = infer_size(x_dst, x_src, size)
    src, dst = edge_index[0], edge_index[1]
    x_i = x_dst.index_select(0, dst)
    x_j = x_src.index_select(0, src)
    edges_new = mlp_forward(torch.cat([x_i, x_j, edge_attr], dim=1), p, prefix, n_extra_layers, activation) + edge_attr
    idx = dst.view(-1, 1).expand_as(edges_new)
    out = edges_new.new_zeros((nd, edges_new.shape[1])).scatter_add_(0, idx, edges_new)
    return out, edges_new


# --------------------------------------------------------------------------------------------------
# GraphTransformerConv with `lin_edge` folded in (round-2 kernel design) -- numpy float64
# --------------------------------------------------------------------------------------------------
def gt_conv_edge_folded_f64(q, k, v, raw, W, b, edge_index, g=None, num_dst: Optional[int] = None):
    """The same conv when `edge_attr` is the block's `lin_edge(raw)` (block.py:497: e_t = W raw_t + b, W [H*C, ed]),
    WITHOUT ever forming an [E, H, C] tensor.  With W_h [C, ed], b_h [C] the rows of head h:

      q_i.e_t      = (W_h^T q_i).raw_t + q_i.b_h            -> per-dst vector  qW_i = W_h^T q_i  (ed values), scalar qb_i
      sum_t a_t e_t = W_h (sum_t a_t raw_t) + b_h sum_t a_t  -> per-dst accumulator R_i = sum_t a_t raw_t (ed values), A_i = sum_t a_t
      g_i.e_t      = (W_h^T g_i).raw_t + g_i.b_h            -> gW_i, gb_i
      dq_i         = (sum_t ds_t k_j + W_h S_i + b_h sum_t ds_t)/sqrt(C),   S_i = sum_t ds_t raw_t
      de_t = a_t g_i + ds_t q_i/sqrt(C) is never formed; what lin_edge's backward needs from it is
      dW_h   = sum_i g_i (x) R_i + q_i (x) S_i / sqrt(C)
      db_h   = sum_i g_i A_i + q_i (sum_t ds_t)/sqrt(C)
      draw_t = sum_h a_t gW_i + ds_t qW_i / sqrt(C)
    Per edge the kernels then touch k_j, v_j and the ed raw values only.  Returns the same keys as `gt_conv_csr_f64`
    plus dW, db, draw (and no `de`)."""
    f = lambda t: np.asarray(t.detach().cpu().to(torch.float64).numpy() if isinstance(t, Tensor) else t, dtype=np.float64)
    q, k, v, raw, W, b = f(q), f(k), f(v), f(raw), f(W), f(b)
    ei = np.asarray(edge_index.cpu().numpy() if isinstance(edge_index, Tensor) else edge_index).astype(np.int64)
    nd = q.shape[0] if num_dst is None else num_dst
    H, C = q.shape[1], q.shape[2]
    ed = raw.shape[1]
    Wh, bh = W.reshape(H, C, ed), b.reshape(H, C)
    src, dst = ei[0], ei[1]
    scale = 1.0 / math.sqrt(C)
    qW = np.einsum("ihc,hcm->ihm", q, Wh)  # [Nd,H,ed]
    qb = np.einsum("ihc,hc->ih", q, bh)
    s = ((q[dst] * k[src]).sum(-1) + (qW[dst] * raw[:, None, :]).sum(-1) + qb[dst]) * scale  # [E,H]
    m = np.full((nd, H), -np.inf)
    np.maximum.at(m, dst, s)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    p = np.exp(s - m_safe[dst])
    l = np.zeros((nd, H))
    np.add.at(l, dst, p)
    a = p / (l[dst] + 1e-16)
    out = np.zeros((nd, H, C))
    np.add.at(out, dst, a[..., None] * v[src])
    R = np.zeros((nd, H, ed))
    np.add.at(R, dst, a[..., None] * raw[:, None, :])
    A = np.zeros((nd, H))
    np.add.at(A, dst, a)
    out = out + np.einsum("ihm,hcm->ihc", R, Wh) + A[..., None] * bh[None]
    res = {"out": out, "lse": m_safe + np.log(l + 1e-16), "alpha": a}
    if g is not None:
        g = f(g)
        gW = np.einsum("ihc,hcm->ihm", g, Wh)
        gb = np.einsum("ihc,hc->ih", g, bh)
        Dl = (g * out).sum(-1)
        gv = (g[dst] * v[src]).sum(-1) + (gW[dst] * raw[:, None, :]).sum(-1) + gb[dst]
        ds = a * (gv - Dl[dst])
        S = np.zeros((nd, H, ed))
        np.add.at(S, dst, ds[..., None] * raw[:, None, :])
        Z = np.zeros((nd, H))
        np.add.at(Z, dst, ds)
        dq = np.zeros_like(q)
        np.add.at(dq, dst, ds[..., None] * k[src])
        dq = (dq + np.einsum("ihm,hcm->ihc", S, Wh) + Z[..., None] * bh[None]) * scale
        dk = np.zeros_like(k)
        np.add.at(dk, src, ds[..., None] * q[dst] * scale)
        dv = np.zeros_like(v)
        np.add.at(dv, src, a[..., None] * g[dst])
        dW = (np.einsum("ihc,ihm->hcm", g, R) + np.einsum("ihc,ihm->hcm", q, S) * scale).reshape(H * C, ed)
        db = (np.einsum("ihc,ih->hc", g, A) + np.einsum("ihc,ih->hc", q, Z) * scale).reshape(H * C)
        draw = (a[..., None] * gW[dst]).sum(1) + (ds[..., None] * qW[dst]).sum(1) * scale
        res.update(dq=dq, dk=dk, dv=dv, dW=dW, db=db, draw=draw, ds=ds, R=R, S=S)
    return res
