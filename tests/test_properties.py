"""Property tests (hypothesis) of the integer host logic: shard sizes, halo plans built from a replicated edge list."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from anemoi_models_b200.distributed.halo import aligned_bounds_from_ranges, build_bipartite_halo_plan
from anemoi_models_b200.distributed.shapes import bounds_from_shapes, tensor_split_sizes
from oracle import sharding as osh


@given(n=st.integers(0, 5000), parts=st.integers(1, 17))
def test_tensor_split_sizes_match_torch(n, parts):
    assert tensor_split_sizes(n, parts) == [x.shape[0] for x in torch.tensor_split(torch.empty(n), parts)]
    assert tensor_split_sizes(n, parts) == osh.tensor_split_sizes(n, parts)


@settings(max_examples=30, deadline=None)
@given(ns=st.integers(1, 60), nd=st.integers(1, 40), e=st.integers(0, 300), parts=st.integers(1, 6), seed=st.integers(0, 10**6))
def test_halo_plans_partition_the_graph_and_agree_across_ranks(ns, nd, e, parts, seed):
    rng = np.random.default_rng(seed)
    ei_np = np.stack([rng.integers(0, ns, e), rng.integers(0, nd, e)]).astype(np.int64)
    ei = torch.from_numpy(ei_np)
    sb = [0] + list(np.cumsum(tensor_split_sizes(ns, parts)))
    db = [0] + list(np.cumsum(tensor_split_sizes(nd, parts)))
    plans = [build_bipartite_halo_plan(ei, [int(x) for x in sb], [int(x) for x in db], r) for r in range(parts)]
    ref_chunks = osh.edges_1hop_chunks((ns, nd), ei_np, parts)
    seen = []
    for r, p in enumerate(plans):
        assert np.array_equal(p.edge_ids.numpy(), ref_chunks[r])  # the reference's 1-hop chunk, bit-exact
        seen.append(p.edge_ids)
        glob = torch.cat([torch.arange(sb[r], sb[r + 1]), p.halo_ids])
        assert torch.equal(glob[p.local_edge_index[0]], ei[0, p.edge_ids])
        assert torch.equal(p.local_edge_index[1] + int(db[r]), ei[1, p.edge_ids])
        assert p.send_counts[r] == 0 and p.recv_counts[r] == 0 and sum(p.recv_counts) == p.n_halo
        off = 0
        for q in range(parts):  # what r sends to q is exactly what q expects from r, in q's halo order
            cnt = p.send_counts[q]
            assert cnt == plans[q].recv_counts[r]
            rows = p.send_idx[off:off + cnt] + int(sb[r])
            qoff = sum(plans[q].recv_counts[:r])
            assert torch.equal(rows, plans[q].halo_ids[qoff:qoff + cnt])
            off += cnt
    assert torch.equal(torch.sort(torch.cat(seen)).values, torch.arange(e))


@settings(max_examples=25, deadline=None)
@given(ns=st.integers(2, 60), nd=st.integers(2, 40), e=st.integers(1, 300), parts=st.integers(2, 8), seed=st.integers(0, 10**6))
def test_peer_push_tables_route_every_row_home(ns, nd, e, parts, seed):
    """Simulate the NVLink push exchange of distributed/peer.py on the CPU: after every rank has pushed, each halo buffer holds
    exactly the rows its plan expects, and after the backward push + per-peer row-add every owner holds the sum of the
    gradients of its rows over all ranks."""
    from anemoi_models_b200.distributed.peer import push_tables

    rng = np.random.default_rng(seed)
    ei = torch.from_numpy(np.stack([rng.integers(0, ns, e), rng.integers(0, nd, e)]).astype(np.int64))
    sb = [0] + [int(x) for x in np.cumsum(tensor_split_sizes(ns, parts))]
    db = [0] + [int(x) for x in np.cumsum(tensor_split_sizes(nd, parts))]
    plans = [build_bipartite_halo_plan(ei, sb, db, r) for r in range(parts)]
    recv_m, send_m = [p.recv_counts for p in plans], [p.send_counts for p in plans]
    x = torch.arange(ns, dtype=torch.float64) * 10 + 1  # row j of the (virtual) full tensor carries the value 10 j + 1
    halo = [torch.full((max(p.n_halo, 1),), -1.0, dtype=torch.float64) for p in plans]
    inbox = [torch.zeros(max(sum(p.send_counts), 1), dtype=torch.float64) for p in plans]
    tables = [push_tables(r, plans[r].send_counts, plans[r].recv_counts, recv_m, send_m) for r in range(parts)]
    for r, p in enumerate(plans):  # forward push
        peer, dst_row, _, _ = tables[r]
        own = x[sb[r]:sb[r + 1]]
        for s_, (q_, d_) in enumerate(zip(peer, dst_row)):
            halo[q_][d_] = own[p.send_idx[s_]]
    for r, p in enumerate(plans):
        assert torch.equal(halo[r][:p.n_halo], x[p.halo_ids])
    grad_halo = [torch.arange(p.n_halo, dtype=torch.float64) + 100 * (r + 1) for r, p in enumerate(plans)]
    for r, p in enumerate(plans):  # backward push
        _, _, owner, inbox_row = tables[r]
        for h_, (q_, d_) in enumerate(zip(owner, inbox_row)):
            inbox[q_][d_] = grad_halo[r][h_]
    expect = torch.zeros(ns, dtype=torch.float64)
    for r, p in enumerate(plans):
        expect.index_add_(0, p.halo_ids, grad_halo[r])
    for r, p in enumerate(plans):  # row-add at the owner
        got = torch.zeros(sb[r + 1] - sb[r], dtype=torch.float64)
        got.index_add_(0, p.send_idx, inbox[r][:sum(p.send_counts)])
        assert torch.equal(got, expect[sb[r]:sb[r + 1]])


@settings(max_examples=60, deadline=None)
@given(ns=st.integers(1, 400), parts=st.integers(1, 9), seed=st.integers(0, 10**6))
def test_aligned_src_bounds_are_a_partition_and_cut_inside_the_shared_zone(ns, parts, seed):
    """Bounds are monotone, cover [0, ns), and on a banded graph (rank r references one window of rows, windows in rank
    order) every cut lies between the next rank's first row and one past the previous ranks' last row: the halo of the
    aligned split never exceeds the overlap of neighbouring windows."""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.integers(0, ns + 1, parts - 1)) if parts > 1 else np.array([], dtype=np.int64)
    edges = np.concatenate([[0], cuts, [ns]])
    ov = int(rng.integers(0, 6))
    lo = [max(0, int(edges[r]) - ov) for r in range(parts)]
    hi = [min(ns - 1, int(edges[r + 1]) - 1 + ov) for r in range(parts)]
    for r in range(parts):  # empty windows become "no edges"
        if edges[r] == edges[r + 1]:
            lo[r], hi[r] = 1, 0
    b = aligned_bounds_from_ranges(lo, hi, ns)
    assert len(b) == parts + 1 and b[0] == 0 and b[-1] == ns and all(b[i] <= b[i + 1] for i in range(parts))
    live = [r for r in range(parts) if lo[r] <= hi[r]]
    for a, c in zip(live[:-1], live[1:]):  # consecutive ranks that have edges
        for cut in b[a + 1:c + 1]:
            assert min(lo[c], hi[a] + 1) <= cut <= max(lo[c], hi[a] + 1)
