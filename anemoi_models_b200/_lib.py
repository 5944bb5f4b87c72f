"""ctypes binding of libanemoi_b200.so (the C ABI declared in include/anemoi_b200.h).

There is no CPU path and no fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from ._build import LIBPATH

AB2_F32, AB2_BF16 = 0, 1
AB2_ERR_INVALID, AB2_ERR_UNSUPPORTED, AB2_ERR_CUDA = 1, 2, 3

_vp, _i64, _i32, _sz, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_float



class Gemm(C.Structure):
    """`ab2_gemm` of include/anemoi_b200.h"""

    _fields_ = [("M", _i64), ("N", _i64), ("K", _i64), ("a", _vp), ("lda", _i64), ("b", _vp), ("ldb", _i64), ("a_mn", C.c_int32),
                ("b_mn", C.c_int32), ("out", _vp * 4), ("ld_out", _i64), ("seg_cols", C.c_int32), ("out_f32", C.c_int32),
                ("bias", _vp), ("row_scale", _vp), ("row_shift", _vp), ("col_vec", _vp), ("pre_out", _vp), ("dact_pre", _vp),
                ("residual", _vp), ("ld_res", _i64), ("res_f32", C.c_int32), ("act", C.c_int32), ("splits", C.c_int32),
                ("reserved", C.c_int32), ("gather_a", _vp), ("gather_a_idx", _vp), ("gather_b", _vp), ("gather_b_idx", _vp),
                ("ld_gather", _i64), ("a_seg", _vp * 3), ("a_seg_len", _i64)]


# name -> (restype, argtypes); must list every symbol of include/anemoi_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "ab2_version": (_i32, []),
    "ab2_last_error": (C.c_char_p, []),
    "ab2_launch_count": (C.c_longlong, []),
    "ab2_csr_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "ab2_csr_build": (_i32, [_vp, _i64, _i64, _i64] + [_vp] * 9 + [_vp, _sz, _vp]),
    "ab2_edge_chunks_workspace_bytes": (_sz, [_i64]),
    "ab2_edge_chunks": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ab2_gtconv_fwd": (_i32, [_vp] * 4 + [_i32] + [_vp] * 3 + [_i64] * 3 + [_i32, _i32, _vp, _vp, _vp]),
    "ab2_gtconv_fold_fwd": (_i32, [_vp] * 5 + [_i32] + [_vp] * 3 + [_i64] * 3 + [_i32, _i32] + [_vp] * 4),
    "ab2_gtconv_fold_bwd_dst": (_i32, [_vp] * 6 + [_i32] + [_vp] * 4 + [_i64] * 3 + [_i32, _i32] + [_vp] * 7),
    "ab2_edge_raw_grad": (_i32, [_vp] * 6 + [_i64, _i64, _i32, _vp, _vp]),
    "ab2_gtconv_variant": (C.c_char_p, [_i32, _i32, _i64, _i64, _i64, _i32, _i32]),
    "ab2_gtconv_fwd_halo": (_i32, [_vp] * 5 + [_i64] + [_vp] + [_i32] + [_vp] * 3 + [_i64] * 3 + [_i32, _i32, _vp, _vp, _vp]),
    "ab2_gtconv_bwd_dst_halo": (_i32, [_vp] * 5 + [_i64] + [_vp] + [_i32] + [_vp] * 4 + [_i64] * 3 + [_i32, _i32] + [_vp] * 5 + [_vp, _sz, _vp]),
    "ab2_gtconv_bwd_src_range_halo": (_i32, [_vp] * 2 + [_i32] + [_vp] * 2 + [_i64] * 4 + [_i32, _i32] + [_vp] * 5 + [_i64, _i64, _vp]),
    "ab2_gtconv_bwd_halo": (_i32, [_vp] * 5 + [_i64] + [_vp] + [_i32] + [_vp] * 6 + [_i64] * 3 + [_i32, _i32] + [_vp] * 9 + [_vp, _sz, _vp]),
    "ab2_gtconv_bwd_workspace_bytes": (_sz, [_i64, _i32]),
    "ab2_gtconv_bwd_dst": (_i32, [_vp] * 4 + [_i32] + [_vp] * 4 + [_i64] * 3 + [_i32, _i32] + [_vp] * 5 + [_vp, _sz, _vp]),
    "ab2_gtconv_bwd_src": (_i32, [_vp] * 2 + [_i32] + [_vp] * 2 + [_i64] * 3 + [_i32, _i32] + [_vp] * 4),
    "ab2_gtconv_bwd_src_range": (_i32, [_vp] * 2 + [_i32] + [_vp] * 2 + [_i64] * 3 + [_i32, _i32] + [_vp] * 3 + [_i64, _i64, _vp]),
    "ab2_gtconv_bwd": (_i32, [_vp] * 4 + [_i32] + [_vp] * 6 + [_i64] * 3 + [_i32, _i32] + [_vp] * 7 + [_vp, _sz, _vp]),
    "ab2_ipc_alloc": (_i32, [_sz, _vp, _vp]),
    "ab2_ipc_open": (_i32, [_vp, _vp]),
    "ab2_ipc_close": (_i32, [_vp]),
    "ab2_ipc_free": (_i32, [_vp]),
    "ab2_memcpy_d2d": (_i32, [_vp, _vp, _sz, _vp]),
    "ab2_peer_push_rows": (_i32, [_vp] * 5 + [_i64, _i32, _vp, _vp, _i32, _vp]),
    "ab2_peer_push_rows_signal": (_i32, [_vp] * 5 + [_i64, _i32, _vp, _vp, _vp, _vp, C.c_uint32, _i32, _i32, _vp]),
    "ab2_peer_wait_flags": (_i32, [_vp, C.c_uint32, _i32, _i32, _vp]),
    "ab2_rows_add": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "ab2_edge_gather_add_act": (_i32, [_vp] * 4 + [_i64] * 3 + [_i32] * 3 + [_vp] * 3),
    "ab2_edge_gather_add_act_bwd": (_i32, [_vp] * 6 + [_i64] * 3 + [_i32] * 3 + [_vp] * 4),
    "ab2_edge_ln_res_segsum": (_i32, [_vp] * 4 + [_f32] + [_vp] * 2 + [_i64] * 2 + [_i32] * 2 + [_vp] * 5),
    "ab2_ln_bwd_parts": (_i32, []),
    "ab2_edge_ln_res_segsum_bwd": (_i32, [_vp] * 7 + [_i64] * 2 + [_i32] * 2 + [_vp] * 3 + [_i32] + [_vp] * 3),
    "ab2_gemm_workspace_bytes": (_sz, [C.POINTER(Gemm)]),
    "ab2_gemm_bf16": (_i32, [C.POINTER(Gemm), _vp, _sz, _vp]),
    "ab2_ln_parts": (_i32, []),
    "ab2_layernorm_fwd": (_i32, [_vp, _i32, _vp, _vp, _f32, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "ab2_layernorm_bwd": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ab2_colsum": (_i32, [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp]),
    "ab2_edge_segment_sums": (_i32, [_vp] * 5 + [_i64] * 3 + [_i32, _i32, _vp, _vp, _vp]),
    "ab2_pad_cast_rows": (_i32, [_vp, _i32, _i64, _i32, _i64, _vp, _i32, _vp]),
    "ab2_unpad_cast_rows": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _i64, _vp]),
    "ab2_gtconv_host_workspace_bytes": (_sz, [_i64] * 3 + [_i32] * 3),
    "ab2_gtconv_fwd_bwd_host_streamed": (_i32, [_vp] * 5 + [_i32] + [_vp] * 6 + [_i64] * 3 + [_i32] * 2 + [_vp] * 5 + [_vp, _i32, _vp, _sz, _vp]),
    "ab2_gtconv_fwd_bwd_host": (_i32, [_vp] * 5 + [_i32] + [_vp] * 6 + [_i64] * 3 + [_i32] * 2 + [_vp] * 5 + [_vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library.  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIBPATH):
                    raise RuntimeError(
                        f"{LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(anemoi_models_b200 has no CPU or PyTorch fallback)"
                    )
                handle = C.CDLL(LIBPATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int) -> None:
    """Translate a non-zero ab2_status into the Python exception the reference path would raise."""
    if rc == 0:
        return
    msg = lib().ab2_last_error().decode("utf-8", "replace")
    if rc in (AB2_ERR_INVALID, AB2_ERR_UNSUPPORTED):
        raise ValueError(f"libanemoi_b200: {msg}")
    raise RuntimeError(f"libanemoi_b200: {msg}")


def ptr(t) -> int:
    """Device (or pinned host) pointer of a tensor, 0 for None."""
    return 0 if t is None else t.data_ptr()


def dtype_code(dtype) -> int:
    import torch

    if dtype == torch.float32:
        return AB2_F32
    if dtype == torch.bfloat16:
        return AB2_BF16
    raise TypeError(f"anemoi_models_b200 supports float32 and bfloat16 tensors, got {dtype}")


def current_stream(device) -> int:
    """cudaStream_t of the calling thread's current stream (autograd worker threads have their own)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream
