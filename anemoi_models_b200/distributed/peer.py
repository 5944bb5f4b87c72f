"""Halo exchange over NVLink peer memory (CUDA IPC), one process per GPU on one node.

`PeerExchange` belongs to a `HaloPlan`: every rank owns one library-allocated buffer (halo planes that peers push the
k / v rows into, and an inbox that peers push the gradients of those rows into, each double-buffered) and maps the
peers' buffers once.  An exchange is then: ONE push kernel (`ab2_peer_push_rows`, plain stores to peer memory through
NVLink, every SM) -> a stream-ordered barrier (a 1-element NCCL all-reduce, ~20 us) -> local use.  Measured against the
NCCL all-to-all it replaces in profiles/r01 (a dst-row-sharded graph sends nearly all of a rank's halo to one neighbour,
which a single NCCL send/recv pair moves at ~270 GB/s).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .. import _lib
from .halo import HaloPlan

LOGGER = logging.getLogger(__name__)


def _ptr_array(ptrs: List[int]):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr


def push_tables(rank: int, send_counts: List[int], recv_counts: List[int], recv_m: List[List[int]], send_m: List[List[int]]):
    """Where this rank's rows land in its peers' buffers (pure index arithmetic, unit-tested on the CPU).

    forward : send row number s (the i-th row of the block destined to peer p) -> row sum(recv_m[p][:rank]) + i of p's
              halo planes (p numbers its halo rows by owner rank, then by ascending id -- the order of the send list);
    backward: halo row number h (the i-th row owned by peer p) -> row sum(send_m[p][:rank]) + i of p's gradient inbox
              (p's inbox mirrors its send list: blocks by destination rank).
    recv_m[p] / send_m[p] are rank p's recv_counts / send_counts.  Returns four int lists."""
    P = len(send_counts)
    peer_of_send, dst_row = [], []
    for p in range(P):
        cnt = send_counts[p]
        if recv_m[p][rank] != cnt:
            raise ValueError("halo plans of the ranks disagree")
        base = sum(recv_m[p][:rank])
        peer_of_send += [p] * cnt
        dst_row += list(range(base, base + cnt))
    owner, inbox_row = [], []
    for p in range(P):
        cnt = recv_counts[p]
        base = sum(send_m[p][:rank])
        owner += [p] * cnt
        inbox_row += list(range(base, base + cnt))
    return peer_of_send, dst_row, owner, inbox_row


class PeerSetupFailed(RuntimeError):
    """Raised on EVERY rank of the group when the peer-memory set-up failed on any of them (the decision is collective)."""


class PeerExchange:
    def __init__(self, plan: HaloPlan, group, row_bytes: int, device: torch.device):
        L = _lib.lib()
        self.plan, self.group, self.row_bytes, self.device = plan, group, int(row_bytes), device
        P, rank = plan.world, plan.rank
        self.P, self.rank = P, rank
        n_send, n_halo = sum(plan.send_counts), plan.n_halo
        info = [None] * P
        dist.all_gather_object(info, (plan.recv_counts, plan.send_counts, n_halo, n_send), group=group)
        recv_m, send_m = [i[0] for i in info], [i[1] for i in info]
        halo_n, send_n = [i[2] for i in info], [i[3] for i in info]
        peer_of_send, dst_row, owner, inbox_row = push_tables(rank, plan.send_counts, plan.recv_counts, recv_m, send_m)
        i32 = dict(dtype=torch.int32, device=device)
        self.peer_of_send = torch.tensor(peer_of_send, **i32)
        self.dst_row_of_send = torch.tensor(dst_row, **i32)
        self.send_idx32 = plan.send_idx.to(torch.int32).contiguous()
        self.send_idx64 = plan.send_idx.to(torch.int64).contiguous()
        self.owner_of_halo = torch.tensor(owner, **i32)
        self.inbox_row_of_halo = torch.tensor(inbox_row, **i32)
        # ---- buffer: [halo A0 | halo B0 | halo A1 | halo B1 | inbox A0 | inbox B0 | inbox A1 | inbox B1]
        def layout(nh, ns):
            hb, ib = max(nh, 1) * self.row_bytes, max(ns, 1) * self.row_bytes
            offs = {}
            o = 0
            for kind, nbytes in (("halo", hb), ("inbox", ib)):
                for parity in (0, 1):
                    for plane in ("a", "b"):
                        offs[(kind, parity, plane)] = o
                        o += (nbytes + 255) // 256 * 256
            # flag words of the barrier protocol (csrc/peer_exchange.cu): one 32-bit slot per peer and exchange kind, written by
            # the peers; and this rank's own CTA counters of the push kernels
            for kind in ("halo", "inbox"):
                offs[("flags", kind)] = o
                o += 256
                offs[("counter", kind)] = o
                o += 256
            return offs, o
        self.offs, total = layout(n_halo, n_send)
        # ---- local phases (may fail on SOME ranks: IPC not permitted, a GPU pair without P2P) are kept apart from the
        # collective ones: every rank walks through the same sequence of collectives whatever happened locally, and the
        # outcome is agreed on with one MIN all-reduce before anything is used.  A rank that raised half-way would
        # otherwise sit in a different collective than its peers (mismatched NCCL calls: a hang, not a fallback).
        self.base = 0
        self.peer_base: List[int] = []
        self._mapped: List[int] = []
        self._closed = False
        err: Optional[Exception] = None
        handle = (C.c_ubyte * 64)()
        try:
            base_ptr = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(L.ab2_ipc_alloc(total, C.byref(base_ptr), handle))
            self.base = base_ptr.value
        except Exception as exc:  # noqa: BLE001 -- reported through the agreed flag below
            err = exc
        handles = [None] * P
        dist.all_gather_object(handles, (bytes(handle), err is None), group=group)
        if err is None and all(h[1] for h in handles):
            try:
                for p in range(P):
                    if p == rank:
                        self.peer_base.append(self.base)
                    else:
                        hp = (C.c_ubyte * 64)(*handles[p][0])
                        mapped = C.c_void_p()
                        with torch.cuda.device(device):
                            _lib.check(L.ab2_ipc_open(hp, C.byref(mapped)))
                        self._mapped.append(mapped.value)
                        self.peer_base.append(mapped.value)
            except Exception as exc:  # noqa: BLE001
                err = exc
        elif err is None:
            err = RuntimeError("a peer could not allocate its exchange buffer")
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int64, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: everyone has mapped everyone before the first push
        if int(ok.item()) == 0:
            self.close()
            raise PeerSetupFailed(str(err) if err is not None else "peer-memory setup failed on another rank")
        peer_offs = [layout(halo_n[p], send_n[p])[0] for p in range(P)]
        self.tables = {}
        for kind in ("halo", "inbox"):
            for parity in (0, 1):
                self.tables[(kind, parity)] = tuple(
                    _ptr_array([self.peer_base[p] + peer_offs[p][(kind, parity, plane)] for p in range(P)]) for plane in ("a", "b"))
        # flag-word barrier: entry p = address of MY slot in rank p's flag array
        self.flag_slots = {kind: _ptr_array([self.peer_base[p] + peer_offs[p][("flags", kind)] + 4 * rank for p in range(P)])
                           for kind in ("halo", "inbox")}
        self.use_flags = os.environ.get("AB2_BARRIER", "flags") != "nccl"  # AB2_BARRIER=nccl: the round-1 all-reduce barrier
        self.token = torch.zeros(1, device=device)
        self.fwd_epoch = 0
        self.bwd_epoch = 0
        # ---- overlap: the exchange runs on a side stream while the conv works on the INTERIOR dst rows (rows none of whose
        # edges references a halo src row); the boundary rows follow once the halo has landed
        # high priority: the push kernel's few CTAs must be scheduled ahead of the tens of thousands of pending conv CTAs
        # of the main stream (measured with AB2_TRACE: at default priority the push only ran once the conv kernel had
        # no more CTAs to launch, i.e. the "overlapped" exchange finished last)
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.overlap = os.environ.get("AB2_OVERLAP", "1") != "0"
        le = plan.local_edge_index
        b_rows = torch.unique(le[1][le[0] >= plan.n_own]).cpu().tolist()
        nd = plan.num_dst_local
        best = (0, 0)
        prev = -1
        for r in b_rows + [nd]:  # longest run of consecutive non-boundary rows
            if r - (prev + 1) > best[1] - best[0]:
                best = (prev + 1, r)
            prev = r
        self.interior = best
        self._ranges = {}

    def close(self) -> None:
        """Unmap the peers' buffers and free the own one (idempotent; called on a failed set-up and when the plan goes away).
        Peers may still hold a mapping of the freed buffer: CUDA keeps the allocation alive until the last mapping closes."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        try:
            L = _lib.lib()
            with torch.cuda.device(self.device):
                for ptr in self._mapped:
                    L.ab2_ipc_close(C.c_void_p(ptr))
                if self.base:
                    L.ab2_ipc_free(C.c_void_p(self.base))
        except Exception:  # noqa: BLE001 -- interpreter shutdown: the driver reclaims everything with the process
            pass
        self._mapped, self.base = [], 0

    def __del__(self):
        self.close()

    def _push(self, kind: str, epoch: int, src_a, src_b, src_row, peer, dst_row, n: int, tables, stream: int) -> None:
        """push rows into the peers' planes and make them visible: flag-word signal + wait (default), or the NCCL barrier."""
        L = _lib.lib()
        ta, tb = tables
        if self.use_flags:
            _lib.check(L.ab2_peer_push_rows_signal(src_a, src_b, src_row, peer, dst_row, n, self.row_bytes, ta, tb,
                                                   self.base + self.offs[("counter", kind)], self.flag_slots[kind], epoch & 0xFFFFFFFF,
                                                   self.P, self.rank, stream))
            _lib.check(L.ab2_peer_wait_flags(self.base + self.offs[("flags", kind)], epoch & 0xFFFFFFFF, self.P, self.rank, stream))
        else:
            _lib.check(L.ab2_peer_push_rows(src_a, src_b, src_row, peer, dst_row, n, self.row_bytes, ta, tb, self.P, stream))
            dist.all_reduce(self.token, group=self.group)

    def _local(self, kind: str, parity: int, plane: str) -> int:
        return self.base + self.offs[(kind, parity, plane)]

    def ranges(self, csr):
        """[(d0, d1, edges)] for the interior block and the boundary blocks before / after it (edge counts from csr.rowptr)."""
        key = id(csr)
        if key not in self._ranges:
            nd = self.plan.num_dst_local
            i_lo, i_hi = self.interior
            pts = [0, i_lo, i_hi, nd]
            rp = csr.rowptr[torch.tensor(pts, device=csr.rowptr.device)].tolist()
            blocks = [(pts[i], pts[i + 1], rp[i + 1] - rp[i]) for i in range(3)]
            self._ranges[key] = {"interior": blocks[1], "boundary": [b for b in (blocks[0], blocks[2]) if b[1] > b[0]]}
        return self._ranges[key]

    def forward_async(self, k: Tensor, v: Tensor):
        """forward() on the side stream; returns (k_halo, v_halo, event to wait for before touching them)."""
        main = torch.cuda.current_stream(self.device)
        outs = tuple(torch.empty((self.plan.n_halo,) + tuple(r.shape[1:]), dtype=r.dtype, device=r.device) for r in (k, v))
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            k_halo, v_halo = self.forward(k, v, outs)
            done = torch.cuda.Event()
            done.record(self.stream)
        return k_halo, v_halo, done

    def backward_async(self, dk_halo: Tensor, dv_halo: Tensor):
        """push the halo gradients on the side stream; returns (buffer parity, event)."""
        L = _lib.lib()
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        b = self.bwd_epoch & 1
        self.bwd_epoch += 1
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            with torch.cuda.device(self.device):
                self._push("inbox", self.bwd_epoch, _lib.ptr(dk_halo), _lib.ptr(dv_halo), 0, _lib.ptr(self.owner_of_halo),
                           _lib.ptr(self.inbox_row_of_halo), self.plan.n_halo, self.tables[("inbox", b)], self.stream.cuda_stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        return b, done

    def backward_finish(self, b: int, dk: Tensor, dv: Tensor) -> None:
        """add what the peers pushed into inbox `b` into dk / dv (current stream; caller has waited for the event)."""
        L = _lib.lib()
        st = _lib.current_stream(self.device)
        D = self.row_bytes // dk.element_size()
        dt = _lib.dtype_code(dk.dtype)
        off = 0
        with torch.cuda.device(self.device):
            for cnt in self.plan.send_counts:
                if cnt:
                    idx_ptr = self.send_idx64.data_ptr() + off * 8
                    for plane, dst in (("a", dk), ("b", dv)):
                        _lib.check(L.ab2_rows_add(_lib.ptr(dst), idx_ptr, self._local("inbox", b, plane) + off * self.row_bytes, cnt,
                                                  D, dt, st))
                off += cnt

    def forward(self, k: Tensor, v: Tensor, outs=None) -> Tuple[Tensor, Tensor]:
        """push the rows of k / v the peers need, receive mine: returns (k_halo, v_halo) [n_halo, ...] (fresh tensors).
        Runs on the current stream."""
        L = _lib.lib()
        plan = self.plan
        b = self.fwd_epoch & 1
        self.fwd_epoch += 1
        st = _lib.current_stream(self.device)
        if outs is None:
            outs = tuple(torch.empty((plan.n_halo,) + tuple(r.shape[1:]), dtype=r.dtype, device=r.device) for r in (k, v))
        with torch.cuda.device(self.device):
            self._push("halo", self.fwd_epoch, _lib.ptr(k), _lib.ptr(v), _lib.ptr(self.send_idx32), _lib.ptr(self.peer_of_send),
                       _lib.ptr(self.dst_row_of_send), self.send_idx32.numel(), self.tables[("halo", b)], st)
            for plane, t in zip(("a", "b"), outs):
                _lib.check(L.ab2_memcpy_d2d(_lib.ptr(t), self._local("halo", b, plane), plan.n_halo * self.row_bytes, st))
        return outs[0], outs[1]

    def backward(self, dk_halo: Tensor, dv_halo: Tensor, dk: Tensor, dv: Tensor) -> None:
        """send the gradients of halo rows home and add what the peers send into dk / dv in place (peer by peer)."""
        L = _lib.lib()
        plan = self.plan
        b = self.bwd_epoch & 1
        self.bwd_epoch += 1
        st = _lib.current_stream(self.device)
        D = dk.numel() // max(dk.shape[0], 1) if dk.shape[0] else self.row_bytes // dk.element_size()
        dt = _lib.dtype_code(dk.dtype)
        with torch.cuda.device(self.device):
            self._push("inbox", self.bwd_epoch, _lib.ptr(dk_halo), _lib.ptr(dv_halo), 0, _lib.ptr(self.owner_of_halo),
                       _lib.ptr(self.inbox_row_of_halo), plan.n_halo, self.tables[("inbox", b)], st)
            off = 0
            for cnt in plan.send_counts:
                if cnt:
                    idx_ptr = self.send_idx64.data_ptr() + off * 8
                    for plane, dst in (("a", dk), ("b", dv)):
                        _lib.check(L.ab2_rows_add(_lib.ptr(dst), idx_ptr, self._local("inbox", b, plane) + off * self.row_bytes, cnt,
                                                  D, dt, st))
                off += cnt


def get_peer_exchange(plan: HaloPlan, group, row_bytes: int, device: torch.device) -> Optional[PeerExchange]:
    """PeerExchange of (plan, row width), created collectively on first use; None -> use the NCCL all-to-all
    (AB2_HALO=nccl, a non-NCCL group, rows that are not a multiple of 16 bytes, or IPC mapping failed on some rank)."""
    cache = plan.__dict__.setdefault("_peer", {})
    key = int(row_bytes)
    if key in cache:
        return cache[key]
    px = None
    usable = (os.environ.get("AB2_HALO", "p2p") != "nccl" and device.type == "cuda" and dist.get_backend(group) == "nccl"
              and row_bytes % 16 == 0 and plan.world <= 16)
    # `usable` depends on the environment of each process: agree on it first, so that either every rank enters the
    # constructor (whose collectives are unconditional, see there) or none does
    ok = torch.tensor([1 if usable else 0], dtype=torch.int64, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        try:
            px = PeerExchange(plan, group, row_bytes, device)
        except PeerSetupFailed as exc:  # raised on every rank alike
            LOGGER.warning("peer-memory halo exchange unavailable (%s); using the NCCL all-to-all", exc)
            px = None
    cache[key] = px
    return px
