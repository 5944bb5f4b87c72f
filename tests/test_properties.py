"""Property tests (hypothesis) of the integer host logic: shard sizes, halo plans built from a replicated edge list."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from anemoi_models_b200.distributed.halo import build_bipartite_halo_plan
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
