"""dst-sorted CSR / src-sorted CSC plan of an `edge_index`, built once per graph on the GPU.

The reference re-derives its gather/scatter plan on every conv call (PyG `propagate` -> `index_select` /
`scatter_add_`, reference layers/conv.py:64,110) although `edge_index` is a constant buffer
(reference layers/mapper.py:144-148, 254).  Here the plan is built by `ab2_csr_build` and cached.
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

_INT_DTYPES = (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64)


def check_edge_index(edge_index) -> None:
    """Same complaints as PyG's MessagePassing._check_input (raised from propagate, reference conv.py:64,110)."""
    if not isinstance(edge_index, Tensor):
        raise ValueError("`MessagePassing.propagate` only supports integer tensors of shape `[2, num_messages]`")
    if edge_index.dtype not in _INT_DTYPES:
        raise ValueError(f"Expected 'edge_index' to be of integer type (got '{edge_index.dtype}')")
    if edge_index.dim() != 2:
        raise ValueError(f"Expected 'edge_index' to be two-dimensional (got {edge_index.dim()} dimensions)")
    if edge_index.size(0) != 2:
        raise ValueError(f"Expected 'edge_index' to have size '2' in the first dimension (got '{edge_index.size(0)}')")


def resolve_size(size, n_src: int, n_dst: int) -> Tuple[int, int]:
    """PyG `_set_size`: fill None entries of size=(Ns, Nd) from the tensors, ValueError on mismatch."""
    ns, nd = (None, None) if size is None else (size[0], size[1])
    if ns is None:
        ns = n_src
    elif ns != n_src:
        raise ValueError(f"Encountered tensor with size {n_src} in dimension 0, but expected size {ns}.")
    if nd is None:
        nd = n_dst
    elif nd != n_dst:
        raise ValueError(f"Encountered tensor with size {n_dst} in dimension 0, but expected size {nd}.")
    return int(ns), int(nd)


class GraphCSR:
    """Device-resident plan: rowptr/col/perm/rowidx (dst-sorted) and colptr/cpos/crow/csr2csc (src-sorted view)."""

    __slots__ = ("num_src", "num_dst", "num_edges", "rowptr", "col", "perm", "rowidx", "colptr", "cpos", "crow", "csr2csc",
                 "perm_is_identity", "edge_index", "device", "_host_meta")

    def __init__(self, edge_index: Tensor, num_src: int, num_dst: int):
        check_edge_index(edge_index)
        if not edge_index.is_cuda:
            raise RuntimeError("anemoi_models_b200 runs on CUDA tensors only (no CPU fallback): edge_index is on the CPU")
        ei = edge_index.to(torch.int64).contiguous()
        E = ei.shape[1]
        dev = ei.device
        L = _lib.lib()
        i32 = dict(dtype=torch.int32, device=dev)
        self.num_src, self.num_dst, self.num_edges, self.device = int(num_src), int(num_dst), int(E), dev
        self.edge_index = ei
        self.rowptr = torch.empty(num_dst + 1, **i32)
        self.col = torch.empty(E, **i32)
        self.perm = torch.empty(E, **i32)
        self.rowidx = torch.empty(E, **i32)
        self.colptr = torch.empty(num_src + 1, **i32)
        self.cpos = torch.empty(E, **i32)
        self.crow = torch.empty((E, 2), **i32)  # (dst, src) of src-sorted position t
        self.csr2csc = torch.empty(E, **i32)
        flags = torch.empty(4, **i32)
        ws_bytes = L.ab2_csr_workspace_bytes(E, num_src, num_dst)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.ab2_csr_build(_lib.ptr(ei), E, num_src, num_dst, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                       _lib.ptr(self.perm), _lib.ptr(self.rowidx), _lib.ptr(self.colptr),
                                       _lib.ptr(self.cpos), _lib.ptr(self.crow), _lib.ptr(self.csr2csc), _lib.ptr(flags),
                                       _lib.ptr(ws), ws_bytes,
                                       _lib.current_stream(dev)))
        f = flags.tolist()  # one-off sync: the plan is cached
        if f[1] != 0:
            raise IndexError(f"edge_index has {f[1]} edge(s) with a node id outside size=({num_src}, {num_dst})")
        self.perm_is_identity = bool(f[0])
        self._host_meta = {}


def tensor_version(t: Tensor) -> int:
    """`t._version`, or -1 for tensors created under `torch.inference_mode()` (they do not track a version counter and
    reading it raises; Lightning's validate / predict loops and inference runners build `edge_index` that way)."""
    return -1 if t.is_inference() else t._version


class TensorKeyedCache:
    """Small LRU of objects derived from an index tensor (+ a hashable extra key).

    Fast path: the very same tensor object with an unchanged `_version` -> no device work at all.  The mapper
    re-creates the expanded `edge_index` with `torch.cat` on every forward (reference mapper.py:254), so on an
    identity miss an entry with the same metadata is reused when `torch.equal` confirms the content (one
    16 B/edge compare instead of a rebuild).  Entries keep a reference to the tensor they were built from, so a
    recycled `data_ptr` can never alias a different graph."""

    def __init__(self, maxsize: int = 16):
        self._items: "OrderedDict[int, tuple]" = OrderedDict()  # serial -> (tensor, version, extra, value)
        self._lock = threading.Lock()
        self._serial = 0
        self._max = maxsize

    def get(self, tensor: Tensor, extra, builder):
        with self._lock:
            entries = list(self._items.items())
        for pass_no in (0, 1):
            for key, (src, ver, ext, value) in reversed(entries):
                if ext != extra or ver != tensor_version(src):
                    continue
                if pass_no == 0:
                    hit = src is tensor
                else:
                    hit = (src.shape == tensor.shape and src.device == tensor.device and src.dtype == tensor.dtype
                           and torch.equal(src, tensor))
                if hit:
                    with self._lock:
                        if key in self._items:
                            self._items.move_to_end(key)
                    return value
        value = builder()
        with self._lock:
            self._serial += 1
            self._items[self._serial] = (tensor, tensor_version(tensor), extra, value)
            while len(self._items) > self._max:
                self._items.popitem(last=False)
        return value

    def clear(self) -> None:
        with self._lock:
            self._items.clear()


_csr_cache = TensorKeyedCache()


def get_csr(edge_index: Tensor, num_src: int, num_dst: int) -> GraphCSR:
    """Cached dst-sorted CSR / src-sorted CSC plan of `edge_index` for a (num_src, num_dst) bipartite graph."""
    check_edge_index(edge_index)
    return _csr_cache.get(edge_index, (int(num_src), int(num_dst)), lambda: GraphCSR(edge_index, num_src, num_dst))


def clear_csr_cache() -> None:
    _csr_cache.clear()
