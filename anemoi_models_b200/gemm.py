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
         residual: Optional[Tensor] = None, splits: int = 1):
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
    ws = None
    need = L.ab2_gemm_workspace_bytes(C.byref(d))
    if need:
        ws = torch.empty(need, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(L.ab2_gemm_bf16(C.byref(d), _p(ws), need, _lib.current_stream(a.device)))
    return out[0] if single else out
