"""CSR / CSC build and dst-chunk edge partition on the GPU: bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import gtconv as og
from oracle import sharding as osh

pytestmark = pytest.mark.gpu


def _check_plan(ei_np, ns, nd):
    from anemoi_models_b200.graph import GraphCSR

    ei = torch.from_numpy(ei_np).cuda()
    plan = GraphCSR(ei, ns, nd)
    rowptr, col, perm = og.csr_build(ei_np, nd)
    assert np.array_equal(plan.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(plan.perm.cpu().numpy(), perm)
    assert np.array_equal(plan.col.cpu().numpy(), col)
    assert np.array_equal(plan.rowidx.cpu().numpy(), ei_np[1][perm].astype(np.int32))
    colptr, pos, row = og.csc_of_csr(rowptr, col, ns)
    assert np.array_equal(plan.colptr.cpu().numpy(), colptr)
    assert np.array_equal(plan.cpos.cpu().numpy(), pos)
    assert np.array_equal(plan.crow.cpu().numpy()[:, 0], row)
    assert np.array_equal(plan.crow.cpu().numpy()[:, 1], col[pos])
    inv = np.empty_like(pos)
    inv[pos] = np.arange(len(pos), dtype=pos.dtype)
    assert np.array_equal(plan.csr2csc.cpu().numpy(), inv)
    assert plan.perm_is_identity == bool(np.array_equal(perm, np.arange(ei_np.shape[1])))
    return plan


@pytest.mark.parametrize("ns,nd,E,seed", [(40, 24, 156, 0), (1000, 300, 20000, 1), (5, 3, 1, 2), (70000, 50000, 400000, 3)])
def test_csr_random_graphs(ns, nd, E, seed):
    rng = np.random.default_rng(seed)
    ei = np.stack([rng.integers(0, ns, E), rng.integers(0, nd, E)]).astype(np.int64)
    _check_plan(ei, ns, nd)


def test_csr_golden_graphs_and_sorted_input():
    z = load_golden("gtconv_bipartite.npz")
    _check_plan(z["edge_index"], 40, 24)
    ei = z["edge_index"][:, np.argsort(z["edge_index"][1], kind="stable")]
    plan = _check_plan(ei, 40, 24)
    assert plan.perm_is_identity


def test_csr_empty_graph_and_isolated_nodes():
    plan = _check_plan(np.zeros((2, 0), dtype=np.int64), 7, 5)
    assert plan.rowptr.tolist() == [0] * 6
    _check_plan(np.array([[0, 0, 6], [4, 4, 4]], dtype=np.int64), 7, 5)


@pytest.mark.parametrize("deg", [33, 100, 5000, 20000])
def test_csr_long_segments(deg):
    """Segments longer than a warp (shared-memory sort) and longer than 8192 (in-place global sort)."""
    rng = np.random.default_rng(deg)
    ns, nd = 3000, 40
    hub_src = rng.integers(0, ns, deg)
    rest = 2000
    ei = np.stack([np.concatenate([hub_src, rng.integers(0, ns, rest)]),
                   np.concatenate([np.full(deg, 17), rng.integers(0, nd, rest)])]).astype(np.int64)
    ei = ei[:, rng.permutation(ei.shape[1])]
    _check_plan(ei, ns, nd)


def test_csr_out_of_range_raises():
    from anemoi_models_b200.graph import GraphCSR

    ei = torch.tensor([[0, 9], [1, 0]]).cuda()
    with pytest.raises(IndexError):
        GraphCSR(ei, 5, 3)
    with pytest.raises(IndexError):
        GraphCSR(torch.tensor([[0, 1], [1, 3]]).cuda(), 5, 3)


def test_csr_cache_reuses_plan_for_equal_content():
    from anemoi_models_b200.graph import clear_csr_cache, get_csr

    clear_csr_cache()
    ei = torch.randint(0, 20, (2, 100)).cuda()
    a = get_csr(ei, 20, 20)
    assert get_csr(ei, 20, 20) is a
    assert get_csr(ei.clone(), 20, 20) is a
    ei2 = ei.clone()
    ei2[0, 0] = (ei2[0, 0] + 1) % 20
    assert get_csr(ei2, 20, 20) is not a


def test_edge_chunks_match_reference_partition():
    from anemoi_models_b200.distributed import sort_edges_1hop_chunks
    from anemoi_models_b200.distributed.khop_edges import edge_chunk_order

    z = load_golden("sharding.npz")
    ns, nd, n = (int(x) for x in z["meta"])
    ei = t(z["bip_edge_index"]).cuda()
    for P in (1, 2, 3, 4, 8):
        order, counts = edge_chunk_order(nd, ei, P)
        assert np.array_equal(order.cpu().numpy(), z[f"bip_ids_P{P}"])
        assert counts == z[f"bip_counts_P{P}"].tolist()
    ei = t(z["one_edge_index"]).cuda()
    ea = torch.arange(ei.shape[1], device="cuda").float().view(-1, 1)
    for P in (1, 2, 4, 5):
        ea_list, ei_list = sort_edges_1hop_chunks(n, ea, ei, P)
        assert np.array_equal(torch.cat(ea_list).view(-1).long().cpu().numpy(), z[f"one_ids_P{P}"])
        assert [a.shape[0] for a in ea_list] == z[f"one_counts_P{P}"].tolist()
        assert torch.equal(torch.cat(ei_list, dim=1), ei[:, torch.cat(ea_list).view(-1).long()])


def test_edge_chunks_large_random():
    from anemoi_models_b200.distributed.khop_edges import edge_chunk_order

    rng = np.random.default_rng(5)
    nd, E = 100000, 1500000
    ei_np = np.stack([rng.integers(0, 50000, E), rng.integers(0, nd, E)]).astype(np.int64)
    order, counts = edge_chunk_order(nd, torch.from_numpy(ei_np).cuda(), 8)
    ids, cnt = osh.edges_1hop_sharding((50000, nd), ei_np, 8)
    assert counts == cnt and np.array_equal(order.cpu().numpy(), ids)
