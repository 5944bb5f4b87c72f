"""The C-ABI shared library: builds, loads, exports every symbol include/anemoi_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "anemoi_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ab2_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ("ab2_csr_build", "ab2_gtconv_fwd", "ab2_gtconv_bwd", "ab2_edge_chunks", "ab2_edge_gather_add_act",
                 "ab2_edge_ln_res_segsum", "ab2_gtconv_fwd_bwd_host", "ab2_last_error"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    from anemoi_models_b200 import _build

    path = _build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    for sym in declared_symbols():
        assert hasattr(handle, sym), f"{sym} declared in anemoi_b200.h but not exported"


def test_ctypes_signatures_cover_the_header():
    from anemoi_models_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    L = _lib.lib()
    assert L.ab2_version() >= 101
    # host-only helpers are callable without a GPU
    assert L.ab2_gtconv_bwd_workspace_bytes(1000, 16) == 1000 * 16 * 8
    assert L.ab2_csr_workspace_bytes(10, 100, 50) > 0
    assert L.ab2_edge_chunks_workspace_bytes(1000) > 8000
    assert L.ab2_gtconv_host_workspace_bytes(10, 5, 20, 2, 8, 0) >= (2 * 5 + 2 * 10 + 20) * 2 * 16 * 4


def test_argument_errors_are_reported_without_a_gpu():
    from anemoi_models_b200 import _lib

    L = _lib.lib()
    rc = L.ab2_gtconv_fwd(0, 0, 0, 0, 7, 0, 0, 0, 1, 1, 1, 1, 8, 0, 0, 0)  # bad dtype
    assert rc == _lib.AB2_ERR_INVALID
    assert b"dtype" in L.ab2_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = L.ab2_csr_build(0, 2**31, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)  # E too large for int32 indices
    assert rc == _lib.AB2_ERR_UNSUPPORTED


def test_gemm_descriptor_errors_without_a_gpu():
    """`ab2_gemm_bf16` validates its descriptor before it touches the device: the ctypes struct matches the header's layout
    (a wrong field offset would turn these into different errors) and bad shapes come back as status codes, not crashes."""
    import ctypes as C

    from anemoi_models_b200 import _lib

    L = _lib.lib()
    assert C.sizeof(_lib.Gemm) == 8 * 3 + 8 * 4 + 8 + 8 * 4 + 8 + 8 + 8 * 4 + 8 * 3 + 8 + 8 + 8 + 8 * 5 + 8 * 3 + 8  # natural alignment, no padding surprises
    d = _lib.Gemm()
    assert L.ab2_gemm_workspace_bytes(C.byref(d)) == 0
    d.M, d.N, d.K = 64, 60, 64  # N % 8 != 0
    d.a, d.b = 16, 16
    d.out[0] = 16
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == _lib.AB2_ERR_UNSUPPORTED and b"multiple of 8" in L.ab2_last_error()
    d.N = 64
    d.a = 0
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == _lib.AB2_ERR_INVALID and b"null operand" in L.ab2_last_error()
    d.a, d.splits, d.K = 16, 4, 1024
    assert L.ab2_gemm_workspace_bytes(C.byref(d)) == 4 * 64 * 64 * 4
    d.bias = 16  # split-K takes a plain single output only
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == _lib.AB2_ERR_INVALID and b"split-K" in L.ab2_last_error()
    d.bias, d.splits = 0, 1
    d.a_seg_len = 40  # segmented A: K-major pieces must be whole 64-element k-blocks, and every piece needs its tensor
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == _lib.AB2_ERR_INVALID and b"a_seg_len" in L.ab2_last_error()
    d.a_seg_len = 512
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == _lib.AB2_ERR_INVALID and b"segment 1 is null" in L.ab2_last_error()
    d.M = 0
    assert L.ab2_gemm_bf16(C.byref(d), None, 0, None) == 0  # nothing to do
    assert L.ab2_ln_parts() > 0
    assert L.ab2_layernorm_fwd(16, 1, 16, 16, 1e-5, 10, 12, 16, 1, None, None, None) == _lib.AB2_ERR_UNSUPPORTED  # D % 8


def test_kernel_dispatch_rules_for_the_model_shapes():
    """`ab2_gtconv_variant` answers from the same rules `run_conv` dispatches with (no GPU needed): the shapes of the AIFS-like
    model (2 KB rows) go to the bulk-copy pipelined kernels, except the src pass at a low out-degree, which goes to the
    warp-cooperative LDG kernel; other row widths stay on the LDG kernels; odd head widths on the generic ones."""
    from anemoi_models_b200 import _lib

    L = _lib.lib()

    def names(dtype, ns, nd, e, h, c):
        return [L.ab2_gtconv_variant(w, dtype, ns, nd, e, h, c).decode().split("<")[0] for w in range(3)]

    bf16, f32 = 1, 0
    assert names(bf16, 542080, 40320, 748256, 16, 64) == ["gtconv_fwd_tma_kernel", "gtconv_bwd_dst_tma_kernel", "gtconv_bwd_src_warp_kernel"]
    assert names(bf16, 40320, 40320, 322560, 16, 64) == ["gtconv_fwd_tma_kernel", "gtconv_bwd_dst_tma_kernel", "gtconv_bwd_src_tma_kernel"]
    assert names(bf16, 40320, 542080, 1626240, 16, 64) == ["gtconv_fwd_tma_kernel", "gtconv_bwd_dst_tma_kernel", "gtconv_bwd_src_tma_kernel"]
    assert names(f32, 40320, 40320, 322560, 16, 32) == ["gtconv_fwd_tma_kernel", "gtconv_bwd_dst_tma_kernel", "gtconv_bwd_src_tma_kernel"]  # 2 KB fp32 rows
    # BASELINE configs[0]: D = 256 fp32 (1 KB rows) -> LDG kernels; low in-degree -> 4-row forward
    assert names(f32, 40320, 10944, 51608, 16, 16) == ["gtconv_fwd_rows_kernel", "gtconv_bwd_dst_kernel", "gtconv_bwd_src_warp_kernel"]
    assert names(f32, 10944, 10944, 87552, 16, 16) == ["gtconv_fwd_kernel", "gtconv_bwd_dst_kernel", "gtconv_bwd_src_warp_kernel"]
    # a row layout where a warp straddles two row groups (tpd = 24) keeps the per-thread src pass
    assert names(f32, 100, 100, 500, 3, 32)[2] == "gtconv_bwd_src_kernel"
    # head width that is not a power-of-two multiple of 16 bytes -> generic kernels
    assert names(f32, 100, 100, 500, 4, 5) == ["gtconv_fwd_generic_kernel", "gtconv_bwd_dst_generic_kernel", "gtconv_bwd_src_generic_kernel"]
