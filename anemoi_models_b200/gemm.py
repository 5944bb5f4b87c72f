"""Dense node-side contractions on the tcgen05 kernel (`ab2_gemm_bf16`, csrc/gemm_tc.cu).

`gemm` is the thin ctypes call; `linear` / `linear_act` are the autograd functions the graph blocks use in place of
`nn.Linear` (reference layers/block.py:491-499, 531-538, 615-633; layers/mlp.py:74-84): forward GEMM with the bias /
activation / residual epilogue, backward = dgrad (activation derivative in its epilogue) + split-K wgrad + a column sum
for the bias.  CUDA bf16 only -- there is no CPU path; callers keep `nn.Linear` for anything else (fp32 parity runs).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _lib

ACT_CODES = {"SiLU": 0, "GELU": 1, "ReLU": 2, "Identity": 3}
ACT_NONE = 3


def _p(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def gemm(a: Tensor, b: Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False, lda: Optional[int] = None,
         ldb: Optional[int] = None, out: Optional[Sequence[Tensor]] = None, seg_cols: int = 0, out_dtype: torch.dtype = torch.bfloat16,
         bias: Optional[Tensor] = None, row_scale: Optional[Tensor] = None, row_shift: Optional[Tensor] = None,
         col_vec: Optional[Tensor] = None, act: int = ACT_NONE, pre_out: Optional[Tensor] = None, dact_pre: Optional[Tensor] = None,
         residual: Optional[Tensor] = None, splits: int = 1, gather=None, a_segs: Optional[Sequence[Tensor]] = None,
         a_seg_len: int = 0):
    """D[M,N] = epilogue(A . B^T) on the tensor cores; see include/anemoi_b200.h (`ab2_gemm`).  a: [M,K] (or [K,M] when a_mn),
    b: [N,K] (or [K,N] when b_mn), bf16, last dim contiguous.  Returns the output tensor (or the list `out` when given)."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("anemoi_models_b200.gemm runs on CUDA tensors only (no CPU fallback)")
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise TypeError("anemoi_models_b200.gemm takes bf16 operands")
    if a.stride(-1) != 1 or b.stride(-1) != 1:
        raise ValueError("gemm operands must be contiguous in their last dimension")
    L = _lib.lib()
    d = _lib.Gemm()
    d.M, d.N, d.K = M, N, K
    d.a, d.lda, d.a_mn = a.data_ptr(), (lda if lda is not None else a.stride(0)), int(a_mn)
    d.b, d.ldb, d.b_mn = b.data_ptr(), (ldb if ldb is not None else b.stride(0)), int(b_mn)
    single = out is None
    if single:
        out = [torch.empty((M, N), dtype=out_dtype, device=a.device)]
    for i, o in enumerate(out):
        if o.dtype != out[0].dtype or o.stride(-1) != 1 or o.stride(0) != out[0].stride(0):
            raise ValueError("gemm outputs must share dtype and row stride")
        d.out[i] = o.data_ptr()
    d.ld_out = out[0].stride(0)
    d.seg_cols = seg_cols
    d.out_f32 = int(out[0].dtype == torch.float32)
    for name, t_ in (("bias", bias), ("row_scale", row_scale), ("row_shift", row_shift), ("col_vec", col_vec)):
        if t_ is not None and (t_.dtype != torch.float32 or not t_.is_contiguous()):
            raise TypeError(f"gemm {name} must be a contiguous fp32 vector")
    d.bias, d.row_scale, d.row_shift, d.col_vec = _p(bias), _p(row_scale), _p(row_shift), _p(col_vec)
    d.act = act
    for name, t_ in (("pre_out", pre_out), ("dact_pre", dact_pre)):
        if t_ is not None and (t_.dtype != torch.bfloat16 or not t_.is_contiguous() or t_.shape != (M, N)):
            raise TypeError(f"gemm {name} must be a contiguous bf16 [M, N] tensor")
    d.pre_out, d.dact_pre = _p(pre_out), _p(dact_pre)
    if residual is not None:
        if residual.stride(-1) != 1 or residual.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("gemm residual must be fp32 or bf16, contiguous in its last dimension")
        d.residual, d.ld_res, d.res_f32 = residual.data_ptr(), residual.stride(0), int(residual.dtype == torch.float32)
    d.splits = splits
    if a_segs:  # A = [a | a_segs...] along K (K-major) or along M (MN-major), each piece a_seg_len wide, never concatenated
        for i, t_ in enumerate(a_segs):
            if t_.dtype != torch.bfloat16 or t_.stride(-1) != 1 or t_.stride(0) != d.lda or t_.shape != a.shape:
                raise TypeError("gemm A segments must be bf16 tensors of one shape and row stride")
            d.a_seg[i] = t_.data_ptr()
        d.a_seg_len = a_seg_len
    if gather is not None:  # ((table_a [Na, N] bf16, idx_a int64 [M]), (table_b, idx_b)): acc += table[idx[m], n], before the activation
        (ta, ia), (tb, ib) = gather
        for t_, i_ in ((ta, ia), (tb, ib)):
            if t_.dtype != torch.bfloat16 or t_.stride(-1) != 1 or t_.stride(0) != ta.stride(0) or i_.dtype != torch.int64 or not i_.is_contiguous():
                raise TypeError("gemm gather tables must be bf16 with a common row stride, indices contiguous int64")
        d.gather_a, d.gather_a_idx, d.gather_b, d.gather_b_idx, d.ld_gather = ta.data_ptr(), ia.data_ptr(), tb.data_ptr(), ib.data_ptr(), ta.stride(0)
    ws = None
    need = L.ab2_gemm_workspace_bytes(C.byref(d))
    if need:
        ws = torch.empty(need, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(L.ab2_gemm_bf16(C.byref(d), _p(ws), need, _lib.current_stream(a.device)))
    return out[0] if single else out


# ---------------------------------------------------------------------------------------------------------------------
# autograd functions used by the graph blocks
# ---------------------------------------------------------------------------------------------------------------------
import os as _os

_TC_OFF = _os.environ.get("AB2_TC", "1") == "0"  # AB2_TC=0: keep every dense contraction on nn.Linear / cuBLASLt (A/B runs)


def tc_applies(x: Tensor, *dims: int) -> bool:
    """The tcgen05 path takes CUDA tensors that are bf16 or that autocast would cast to bf16, with GEMM dims that are
    multiples of 8; everything else (fp32 parity runs, CPU, odd widths) stays on nn.Linear."""
    if _TC_OFF or not x.is_cuda or any(d % 8 != 0 or d <= 0 for d in dims):
        return False
    if x.dtype == torch.bfloat16:
        return True
    return x.dtype == torch.float32 and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else t.detach().float().contiguous()


def _bf16_weight(w: Tensor) -> Tensor:
    """bf16 operand copy of a weight.  (Caching it per parameter `_version` would save the re-cast in the checkpoint recompute,
    ~1.5 ms of the AIFS-like step, but a weight updated through `.data` does not bump the version: not worth a stale weight.)"""
    return w.detach().to(torch.bfloat16).contiguous()


def colsum(a: Tensor) -> Tensor:
    """fp32 column sums of a [M, N] matrix (bias gradients), deterministic two-stage reduction."""
    L = _lib.lib()
    M, N = a.shape
    parts = L.ab2_ln_parts()
    partial = torch.empty((parts, N), dtype=torch.float32, device=a.device)
    out = torch.empty(N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(L.ab2_colsum(a.data_ptr(), _lib.dtype_code(a.dtype), M, N, a.stride(0), partial.data_ptr(), out.data_ptr(),
                                _lib.current_stream(a.device)))
    return out


# the two kernel calls are module-level functions so that the CPU tests can check the autograd glue against a torch emulation
def _ln_fwd_kernel(x2: Tensor, gamma: Tensor, beta: Tensor, eps: float, out_dtype: torch.dtype):
    L = _lib.lib()
    M, D = x2.shape
    y = torch.empty((M, D), dtype=out_dtype, device=x2.device)
    mean = torch.empty(M, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        _lib.check(L.ab2_layernorm_fwd(x2.data_ptr(), _lib.dtype_code(x2.dtype), gamma.data_ptr(), beta.data_ptr(), eps, M, D,
                                       y.data_ptr(), _lib.dtype_code(out_dtype), mean.data_ptr(), rstd.data_ptr(),
                                       _lib.current_stream(x2.device)))
    return y, mean, rstd


def _ln_bwd_kernel(g2: Tensor, x2: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, add: Optional[Tensor] = None):
    """dx = LN-backward(g) (+ add, the gradient that reached x through a residual branch: one pass instead of an add kernel)"""
    L = _lib.lib()
    M, D = x2.shape
    dx = torch.empty_like(x2)
    parts = L.ab2_ln_parts()
    partial = torch.empty((parts + 1) * 2 * D, dtype=torch.float32, device=x2.device)
    dgamma = torch.empty(D, dtype=torch.float32, device=x2.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        _lib.check(L.ab2_layernorm_bwd(g2.data_ptr(), _lib.dtype_code(g2.dtype), x2.data_ptr(), _lib.dtype_code(x2.dtype),
                                       gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, _p(add), dx.data_ptr(),
                                       partial.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _lib.current_stream(x2.device)))
    return dx, dgamma, dbeta


class _LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim, output directly in the dtype the next GEMM consumes (reference block.py:487-489, 611,
    349: LayerNorm in fp32 under autocast + the cast in front of each nn.Linear)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, eps: float, out_dtype: torch.dtype, fork: bool = False):
        ctx.set_materialize_grads(False)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        gamma, beta = _f32(weight), _f32(bias)
        y, mean, rstd = _ln_fwd_kernel(x2, gamma, beta, float(eps), out_dtype)
        ctx.save_for_backward(x2, gamma, mean, rstd)
        ctx.shape, ctx.pdt = tuple(x.shape), (weight.dtype, bias.dtype)
        if fork:  # (LN(x), x): the skip connection leaves through this node too, so both gradients meet in ONE backward kernel
            return y.view(ctx.shape), x.view_as(x)
        return y.view(ctx.shape)

    @staticmethod
    def backward(ctx, g: Optional[Tensor], g_skip: Optional[Tensor] = None):
        x2, gamma, mean, rstd = ctx.saved_tensors
        M, D = x2.shape
        if g is None:  # only the skip branch reached the loss
            return g_skip, None, None, None, None, None
        g2 = g.reshape(M, D).contiguous()
        if g2.dtype not in (torch.float32, torch.bfloat16):
            g2 = g2.float()
        add = None
        if g_skip is not None:
            add = g_skip.reshape(M, D)
            if add.dtype != x2.dtype or not add.is_contiguous():
                add = add.to(x2.dtype).contiguous()
        dx, dgamma, dbeta = _ln_bwd_kernel(g2, x2, gamma, mean, rstd, add)
        return dx.view(ctx.shape), dgamma.to(ctx.pdt[0]), dbeta.to(ctx.pdt[1]), None, None, None


def layer_norm(x: Tensor, ln: torch.nn.LayerNorm, out_dtype: torch.dtype = torch.bfloat16) -> Tensor:
    """`ln(x)` written in `out_dtype` (bf16: what the tensor-core GEMM behind it reads)."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return _LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps, out_dtype)


def layer_norm_fork(x: Tensor, ln: torch.nn.LayerNorm, out_dtype: torch.dtype = torch.bfloat16):
    """(`ln(x)`, x) for the pre-norm residual pattern `f(ln(x)) + x` (reference block.py:487-537, 611-633): take the skip
    connection from the second result and its gradient is added inside the LayerNorm backward kernel."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return _LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps, out_dtype, True)


def wgrad_splits(N: int, K: int, M: int, slots: int = 74) -> int:
    """Split count of the wgrad contraction (over the M rows).  dW has only (N/256) x (K/256) output tiles -- 16 for a
    1024 x 1024 weight -- so the contraction is cut into `s` pieces to give every CTA pair work, and `s` is chosen so that the
    tile count fills whole waves of the `slots` resident CTA pairs (160 tiles = 2.16 waves cost three rounds: measured 72 % of
    the one-wave rate), with at least 8 k-blocks per piece and no empty piece."""
    pairs = ((N + 255) // 256) * ((K + 255) // 256 if K > 128 else 1)
    kb = (M + 63) // 64
    best, best_eff = 1, 0.0
    for s_ in range(1, 65):
        if s_ > 1 and (kb // s_ < 8 or ((kb + s_ - 1) // s_) * (s_ - 1) >= kb):
            continue
        tiles = pairs * s_
        eff = tiles / (((tiles + slots - 1) // slots) * slots)
        if eff > best_eff + 0.04:  # prefer fewer splits (fewer fp32 partials) unless the gain is real
            best, best_eff = s_, eff
    return best


class _LinearFn(torch.autograd.Function):
    """y = op(x) W^T + b (+ residual), op(x) = x, or act(pre) when the producer handed over its pre-activation `pre` together
    with h = act(pre) (`x` is then h).  With `act_out` the function returns (y, act(y)): the activation is applied in the
    GEMM epilogue and the consumer differentiates through it in ITS dgrad epilogue (`pre` argument), so no elementwise
    activation / activation-gradient pass exists.  Backward: dgrad (B operand MN-major, activation derivative fused),
    split-K wgrad (both operands MN-major), deterministic column sum for the bias."""

    @staticmethod
    def forward(ctx, x: Tensor, pre: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor],
                act_in: int, act_out: int):
        ctx.set_materialize_grads(False)  # the (non-differentiable) activation output would otherwise get a zero-filled [M, N] gradient
        shape = tuple(x.shape)
        x2 = x.reshape(-1, shape[-1])
        if x2.dtype != torch.bfloat16 or x2.stride(-1) != 1 or (x2.stride(0) % 8) != 0 or (x2.data_ptr() % 16) != 0:
            x2 = x2.to(torch.bfloat16).contiguous()
        M, K = x2.shape
        N = weight.shape[0]
        w = _bf16_weight(weight)
        b = _f32(bias)
        res2 = None
        if residual is not None:
            res2 = residual.reshape(M, N)
            if res2.dtype not in (torch.float32, torch.bfloat16) or res2.stride(-1) != 1:
                res2 = res2.float().contiguous()
        out_dtype = res2.dtype if res2 is not None else torch.bfloat16
        y_pre = None
        if act_out != ACT_NONE:
            y_pre = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            y = gemm(x2, w, M, N, K, bias=b, act=act_out, pre_out=y_pre, residual=res2, out_dtype=out_dtype)
        else:
            y = gemm(x2, w, M, N, K, bias=b, residual=res2, out_dtype=out_dtype)
        ctx.save_for_backward(x2, pre.reshape(M, K) if pre is not None else None, w)
        ctx.meta = (shape, act_in, weight.dtype, None if bias is None else bias.dtype, None if residual is None else residual.dtype,
                    None if residual is None else tuple(residual.shape))
        oshape = shape[:-1] + (N,)
        if act_out != ACT_NONE:
            h = y.view(oshape)
            ctx.mark_non_differentiable(h)
            return y_pre.view(oshape), h
        return y.view(oshape)

    @staticmethod
    def backward(ctx, g: Tensor, _g_h=None):
        if g is None:
            return None, None, None, None, None, None, None
        x2, pre, w = ctx.saved_tensors
        shape, act_in, wdt, bdt, rdt, rshape = ctx.meta
        M, K = x2.shape
        N = w.shape[0]
        need = ctx.needs_input_grad
        g_in = g.reshape(M, N)
        g2 = g_in
        if g2.dtype != torch.bfloat16 or g2.stride(-1) != 1 or (g2.stride(0) % 8) != 0 or (g2.data_ptr() % 16) != 0:
            g2 = g2.to(torch.bfloat16).contiguous()
        dx = dpre = dw = db = dres = None
        if need[0] or need[1]:
            # d op(x) = g W : A = g [M, N] (K-major, contraction N), B = W stored [N, K] = [contraction, out] -> MN-major
            if pre is not None:
                dpre = gemm(g2, w, M, K, N, b_mn=True, dact_pre=pre.contiguous(), act=act_in).view(shape)
            else:
                dx = gemm(g2, w, M, K, N, b_mn=True).view(shape)
        if need[2]:
            # dW [N, K] = g^T x : A = g stored [M, N] = [contraction, out rows] (MN-major), B = x stored [M, K] (MN-major)
            dw = gemm(g2, x2, N, K, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=wgrad_splits(N, K, M)).to(wdt)
        if need[3]:
            db = colsum(g2).to(bdt)
        if need[4]:
            dres = g_in.to(rdt).reshape(rshape)
        return dx, dpre, dw, db, dres, None, None


class _MultiLinearFn(torch.autograd.Function):
    """y_i = x W_i^T + b_i for 2-4 Linear layers of ONE shape reading the same x -- the block's q | k | v | self projections
    (reference block.py:491-499, 613-620).  Forward: one GEMM over the stacked weights, each y_i written to its own tensor
    (column segments of the epilogue).  Backward: one dgrad whose A operand is [dy_1 | ... | dy_n] read where autograd left the
    pieces (segmented tensor maps, no concatenation, no add kernels between four partial dx), one split-K wgrad over the same
    segmented operand, column sums for the biases."""

    @staticmethod
    def forward(ctx, x: Tensor, n: int, *wb):
        weights, biases = wb[:n], wb[n:]
        shape = tuple(x.shape)
        x2 = x.reshape(-1, shape[-1])
        if x2.dtype != torch.bfloat16 or x2.stride(-1) != 1 or (x2.stride(0) % 8) != 0 or (x2.data_ptr() % 16) != 0:
            x2 = x2.to(torch.bfloat16).contiguous()
        M, K = x2.shape
        N = weights[0].shape[0]
        wcat = torch.cat([_bf16_weight(w) for w in weights], dim=0)
        bcat = None if biases[0] is None else torch.cat([b.detach().float() for b in biases])
        outs = [torch.empty((M, N), dtype=torch.bfloat16, device=x.device) for _ in range(n)]
        gemm(x2, wcat, M, n * N, K, out=outs, seg_cols=N, bias=bcat)
        ctx.save_for_backward(x2, wcat)
        ctx.meta = (shape, n, N, [w.dtype for w in weights], [None if b is None else b.dtype for b in biases])
        return tuple(o.view(shape[:-1] + (N,)) for o in outs)

    @staticmethod
    def backward(ctx, *gs):
        x2, wcat = ctx.saved_tensors
        shape, n, N, wdts, bdts = ctx.meta
        M, K = x2.shape
        need = ctx.needs_input_grad
        g2 = []
        for g in gs:  # materialised by autograd (zeros for an unused output)
            g_ = g.reshape(M, N)
            if g_.dtype != torch.bfloat16 or not g_.is_contiguous() or (g_.data_ptr() % 16) != 0:
                g_ = g_.to(torch.bfloat16).contiguous()
            g2.append(g_)
        dx = None
        if need[0]:  # dx = [dy_1 | ... | dy_n] Wcat : A segmented along the contraction (n * N), B = Wcat stored [n * N, K] (MN-major)
            dx = gemm(g2[0], wcat, M, K, n * N, b_mn=True, a_segs=g2[1:], a_seg_len=N).view(shape)
        dws = [None] * n
        if any(need[2:2 + n]):  # dWcat [n * N, K] = [dy_1 | ... | dy_n]^T x : A MN-major, segmented along its rows
            dwcat = gemm(g2[0], x2, n * N, K, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=wgrad_splits(n * N, K, M),
                         a_segs=g2[1:], a_seg_len=N)
            dws = [dwcat[i * N:(i + 1) * N].to(wdts[i]) if need[2 + i] else None for i in range(n)]
        dbs = [colsum(g2[i]).to(bdts[i]) if (bdts[i] is not None and need[2 + n + i]) else None for i in range(n)]
        return (dx, None, *dws, *dbs)


class _PadCastFn(torch.autograd.Function):
    """[rows, k] raw features (fp32 / bf16) -> bf16 [rows, kp] zero-padded rows for the `lin_edge` GEMM, one kernel each way."""

    @staticmethod
    def forward(ctx, x: Tensor, kp: int) -> Tensor:
        rows, k = x.shape
        if x.stride(-1) != 1 or x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float().contiguous()
        out = torch.empty((rows, kp), dtype=torch.bfloat16, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            _lib.check(L.ab2_pad_cast_rows(x.data_ptr(), _lib.dtype_code(x.dtype), rows, k, x.stride(0), out.data_ptr(), kp,
                                           _lib.current_stream(x.device)))
        ctx.meta = (rows, k, kp, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        rows, k, kp, dt = ctx.meta
        g2 = g if (g.dtype == torch.bfloat16 and g.is_contiguous()) else g.to(torch.bfloat16).contiguous()
        dx = torch.empty((rows, k), dtype=dt, device=g.device)
        L = _lib.lib()
        with torch.cuda.device(g.device):
            _lib.check(L.ab2_unpad_cast_rows(g2.data_ptr(), rows, kp, dx.data_ptr(), _lib.dtype_code(dt), k, k,
                                             _lib.current_stream(g.device)))
        return dx, None


def pad_cast(x: Tensor, multiple: int = 16) -> Tensor:
    """bf16 copy of the rows of `x` [rows, k], zero-padded to a multiple of `multiple` columns (differentiable)."""
    kp = ((x.shape[-1] + multiple - 1) // multiple) * multiple
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return _PadCastFn.apply(x, kp)


def linear_multi(x: Tensor, lins: Sequence[torch.nn.Linear]):
    """[lin(x) for lin in lins] for 2-4 Linear layers of one shape (out_features a multiple of 256): one forward GEMM, one
    dgrad, one wgrad (`_MultiLinearFn`); other shapes fall back to one GEMM per layer."""
    n = len(lins)
    N, K = lins[0].weight.shape
    same = all(tuple(l.weight.shape) == (N, K) and (l.bias is None) == (lins[0].bias is None) for l in lins)
    if not (2 <= n <= 4 and same and N % 256 == 0):
        return [linear(x, l) for l in lins]
    return list(_MultiLinearFn.apply(x, n, *[l.weight for l in lins], *[l.bias for l in lins]))


def linear(x: Tensor, lin: torch.nn.Linear, residual: Optional[Tensor] = None, act_out: int = ACT_NONE):
    """`lin(x)` (+ residual) on the tensor-core kernel; with `act_out` returns (pre-activation, activation) for `act_linear`."""
    return _LinearFn.apply(x, None, lin.weight, lin.bias, residual, ACT_NONE, act_out)


def linear_wb(x: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor] = None, act_out: int = ACT_NONE):
    """`F.linear(x, weight, bias)` (+ residual) for a weight that is a slice of a parameter (GraphConv's split first layer)."""
    return _LinearFn.apply(x, None, weight, bias, residual, ACT_NONE, act_out)


def act_linear(pre: Tensor, h: Tensor, lin: torch.nn.Linear, act: int, residual: Optional[Tensor] = None, act_out: int = ACT_NONE):
    """`lin(act(pre))` (+ residual) where h = act(pre) came out of the previous GEMM's epilogue."""
    return _LinearFn.apply(h, pre, lin.weight, lin.bias, residual, act, act_out)


def _segment_sums_kernel(g2: Tensor, plan, want_dst: bool, want_src: bool):
    """(sum of g over the edges into every dst, sum over the edges out of every src), bf16 [E, N] in original edge order"""
    L = _lib.lib()
    E, N = g2.shape
    dpi = torch.empty((plan.num_dst, N), dtype=torch.bfloat16, device=g2.device) if want_dst else None
    dpj = torch.empty((plan.num_src, N), dtype=torch.bfloat16, device=g2.device) if want_src else None
    with torch.cuda.device(g2.device):
        _lib.check(L.ab2_edge_segment_sums(g2.data_ptr(), plan.rowptr.data_ptr(), plan.perm.data_ptr(), plan.colptr.data_ptr(),
                                           plan.cpos.data_ptr(), E, plan.num_src, plan.num_dst, N, _lib.AB2_BF16, _p(dpi), _p(dpj),
                                           _lib.current_stream(g2.device)))
    return dpi, dpj


class _EdgeFirstLayerFn(torch.autograd.Function):
    """First layer of GraphConv's edge MLP on the split weight (reference conv.py:69, mlp.py:74):
    pre[t] = e[t] We^T + pi[dst_t] + pj[src_t], h = act(pre) -- ONE GEMM over the edges whose epilogue gathers the two node-side
    rows (pi, pj are L2-resident node tables), applies the activation and keeps the pre-activation.  Returns (pre, h), h not
    differentiable (the consumer differentiates through the activation in its dgrad epilogue).
    Backward: de = dpre We (dgrad), dWe = dpre^T e (split-K wgrad), dpi / dpj = CSR / CSC segment sums of dpre."""

    @staticmethod
    def forward(ctx, e: Tensor, pi: Tensor, pj: Tensor, w_e: Tensor, plan, act: int):
        ctx.set_materialize_grads(False)
        E, K = e.shape
        N = w_e.shape[0]
        e2 = e if (e.dtype == torch.bfloat16 and e.is_contiguous()) else e.to(torch.bfloat16).contiguous()
        w = w_e.detach().to(torch.bfloat16).contiguous()
        pi2, pj2 = pi.to(torch.bfloat16).contiguous(), pj.to(torch.bfloat16).contiguous()
        pre = torch.empty((E, N), dtype=torch.bfloat16, device=e.device)
        ei = plan.edge_index
        h = gemm(e2, w, E, N, K, act=act, pre_out=pre, gather=((pi2, ei[1]), (pj2, ei[0])))
        ctx.save_for_backward(e2, w)
        ctx.plan, ctx.dts = plan, (e.dtype, pi.dtype, pj.dtype, w_e.dtype)
        ctx.mark_non_differentiable(h)
        return pre, h

    @staticmethod
    def backward(ctx, g: Tensor, _gh=None):
        if g is None:
            return None, None, None, None, None, None
        e2, w = ctx.saved_tensors
        plan = ctx.plan
        L = _lib.lib()
        E, K = e2.shape
        N = w.shape[0]
        g2 = g if (g.dtype == torch.bfloat16 and g.is_contiguous()) else g.to(torch.bfloat16).contiguous()
        need = ctx.needs_input_grad
        de = gemm(g2, w, E, K, N, b_mn=True).to(ctx.dts[0]) if need[0] else None
        dw = gemm(g2, e2, N, K, E, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=wgrad_splits(N, K, E)).to(ctx.dts[3]) if need[3] else None
        dpi, dpj = _segment_sums_kernel(g2, plan, need[1], need[2]) if (need[1] or need[2]) else (None, None)
        return (de, None if dpi is None else dpi.to(ctx.dts[1]), None if dpj is None else dpj.to(ctx.dts[2]), dw, None, None)


def edge_first_layer(e: Tensor, pi: Tensor, pj: Tensor, w_e: Tensor, plan, act: int):
    return _EdgeFirstLayerFn.apply(e, pi, pj, w_e, plan, act)


def mlp_tail(pre: Tensor, h: Tensor, linears, act: int) -> Tensor:
    """The Linear layers that follow an activation already applied in a GEMM epilogue: (Linear, act)* Linear."""
    for lin in linears[:-1]:
        pre, h = act_linear(pre, h, lin, act, act_out=act)
    return act_linear(pre, h, linears[-1], act)


def mlp_forward(x: Tensor, linears, act: int) -> Tensor:
    """Linear, act, (Linear, act)*, Linear (reference mlp.py:74-80) with every activation in a GEMM epilogue."""
    pre, h = linear(x, linears[0], act_out=act)
    return mlp_tail(pre, h, linears[1:], act)
