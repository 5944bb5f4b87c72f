"""The oracle restatements against the golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, t
from oracle import blocks as oblocks
from oracle import gtconv as og
from oracle import sharding as osh

GT_CASES = ["gtconv_bipartite.npz", "gtconv_c64.npz", "gtconv_c16_h16.npz", "gtconv_biglogit.npz", "gtconv_oddc.npz",
            "gtconv_kat3.npz"]


@pytest.mark.parametrize("name", GT_CASES)
def test_gtconv_unfused_matches_reference_bitwise(name):
    torch.set_num_threads(1)
    z = load_golden(name)
    size = tuple(int(x) for x in z["size"])
    r = og.gt_conv_unfused_fwd_bwd(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), t(z["g"]), size)
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(r[key], t(z[key])) < 1e-6, key


@pytest.mark.parametrize("name", GT_CASES)
def test_gtconv_csr_formulas_match_reference(name):
    z = load_golden(name)
    r = og.gt_conv_csr_f64(z["q"], z["k"], z["v"], z["e"], z["edge_index"], z["g"])
    for key in ("out", "dq", "dk", "dv", "de"):
        assert rel_err(torch.from_numpy(r[key]), t(z[key])) < 2e-5, key
    loops = og.gt_conv_loops_f64(z["q"], z["k"], z["v"], z["e"], z["edge_index"])
    assert np.abs(loops - r["out"]).max() < 1e-10


def test_kat3_by_hand():
    z = load_golden("gtconv_kat3.npz")
    out = og.gt_conv_unfused(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), (2, 3))
    assert torch.allclose(out, t(z["expect"]), atol=1e-6)
    assert torch.equal(out[2], torch.zeros_like(out[2]))  # isolated dst -> zero row


def test_size_mismatch_raises():
    z = load_golden("gtconv_kat3.npz")
    with pytest.raises(ValueError):
        og.gt_conv_unfused(t(z["q"]), t(z["k"]), t(z["v"]), t(z["e"]), t(z["edge_index"]), (5, 3))


def test_csr_build_is_stable_sort():
    z = load_golden("gtconv_bipartite.npz")
    ei = z["edge_index"]
    rowptr, col, perm = og.csr_build(ei, 24)
    order = torch.sort(torch.from_numpy(ei[1]), stable=True).indices.numpy()
    assert np.array_equal(perm, order)
    assert np.array_equal(col, ei[0][order])
    assert rowptr[0] == 0 and rowptr[-1] == ei.shape[1]
    assert np.array_equal(np.diff(rowptr), np.bincount(ei[1], minlength=24))
    colptr, pos, row = og.csc_of_csr(rowptr, col, 40)
    assert np.array_equal(np.diff(colptr), np.bincount(ei[0], minlength=40))
    assert np.array_equal(col[pos], np.sort(col, kind="stable"))
    assert np.array_equal(row, ei[1][perm][pos])


def _params(z, prefix="p."):
    return {k[len(prefix):]: t(v) for k, v in z.items() if k.startswith(prefix)}


def test_graphconv_matches_reference():
    torch.set_num_threads(1)
    z = load_golden("graphconv_bipartite.npz")
    p = _params(z)
    out, en = og.graph_conv_unfused((t(z["xs"]), t(z["xd"])), t(z["e"]), t(z["edge_index"]), p, "edge_mlp.",
                                    size=tuple(int(x) for x in z["size"]))
    assert rel_err(out, t(z["out"])) < 1e-6 and rel_err(en, t(z["edges_new"])) < 1e-6
    z = load_golden("graphconv_single.npz")
    out, en = og.graph_conv_unfused(t(z["x"]), t(z["e"]), t(z["edge_index"]), _params(z), "edge_mlp.")
    assert rel_err(out, t(z["out"])) < 1e-6 and rel_err(en, t(z["edges_new"])) < 1e-6


def test_blocks_match_reference():
    torch.set_num_threads(1)
    z = load_golden("block_gt_mapper.npz")
    ns, nd, D, H, ed, hid = (int(x) for x in z["meta"])
    (_, dst), _ = oblocks.gt_mapper_block(_params(z), (t(z["xs"]), t(z["xd"])), t(z["ea"]), t(z["edge_index"]), H, (ns, nd))
    assert rel_err(dst, t(z["dst_new"])) < 1e-6
    z = load_golden("block_gt_processor.npz")
    n, _, D, H, ed, hid = (int(x) for x in z["meta"])
    nodes, _ = oblocks.gt_processor_block(_params(z), t(z["x"]), t(z["ea"]), t(z["edge_index"]), H)
    assert rel_err(nodes, t(z["nodes_new"])) < 1e-6
    z = load_golden("block_graphconv_processor.npz")
    nodes, edges = oblocks.graphconv_processor_block(_params(z), t(z["x"]), t(z["e"]), t(z["edge_index"]))
    assert rel_err(nodes, t(z["nodes_new"])) < 1e-6 and rel_err(edges, t(z["edges_new"])) < 1e-6
    z = load_golden("block_graphconv_mapper.npz")
    ns, nd, D = (int(x) for x in z["meta"])
    (s, d), e = oblocks.graphconv_mapper_block(_params(z), (t(z["xs"]), t(z["xd"])), t(z["e"]), t(z["edge_index"]), (ns, nd))
    assert rel_err(s, t(z["src_new"])) < 1e-6 and rel_err(d, t(z["dst_new"])) < 1e-6 and rel_err(e, t(z["edges_new"])) < 1e-6


def test_sharding_bit_exact():
    z = load_golden("sharding.npz")
    ns, nd, n = (int(x) for x in z["meta"])
    for key, (rows, P) in {"shards_10_3": (10, 3), "shards_40320_8": (40320, 8), "shards_7_8": (7, 8),
                           "shards_542080_8": (542080, 8)}.items():
        assert np.array_equal(np.array(osh.shape_shards((rows, 1), 0, P)), z[key])
    for P in (1, 2, 3, 4, 8):
        ids, counts = osh.edges_1hop_sharding((ns, nd), z["bip_edge_index"], P)
        assert np.array_equal(ids, z[f"bip_ids_P{P}"]) and np.array_equal(np.array(counts), z[f"bip_counts_P{P}"])
    for P in (1, 2, 4, 5):
        ids, counts = osh.edges_1hop_sharding(n, z["one_edge_index"], P)
        assert np.array_equal(ids, z[f"one_ids_P{P}"]) and np.array_equal(np.array(counts), z[f"one_counts_P{P}"])
    assert osh.change_channels([[5, 3], [4, 3]], 7) == [[5, 7], [4, 7]] and osh.change_channels([], 7) == []
    ei = osh.expand_edges(np.array([[0, 1], [2, 0]]), 4, 3, 2)
    assert np.array_equal(ei, np.array([[0, 1, 4, 5], [2, 0, 5, 3]]))


def test_halo_plan_covers_every_needed_row():
    z = load_golden("sharding.npz")
    ns, nd, _ = (int(x) for x in z["meta"])
    ei = z["bip_edge_index"]
    for P in (2, 4):
        plans = osh.halo_plan(ei, ns, nd, P)
        assert sum(len(p["edge_ids"]) for p in plans) == ei.shape[1]
        for r, p in enumerate(plans):
            assert np.array_equal(np.unique(ei[0, p["edge_ids"]]), p["needed_src"])
            assert sum(len(v) for v in p["recv_from"].values()) == len(p["needed_src"])


def test_edge_folded_formulation_equals_the_reference_op_sequence():
    """oracle.gt_conv_edge_folded_f64 (lin_edge folded into the conv, no [E,H,C] tensor) against the reference's op sequence
    `conv(q, k, v, lin_edge(raw))` with autograd, float64: outputs and every gradient incl. lin_edge's weight, bias and input."""
    import numpy as np
    import torch

    from oracle import gtconv as og

    gen = torch.Generator().manual_seed(5)
    ns, nd, E, H, C, ed = 37, 19, 160, 4, 8, 11
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd - 2, (E,), generator=gen)])  # 2 isolated dst rows
    q, k, v = (torch.randn(n, H, C, generator=gen, dtype=torch.float64) for n in (nd, ns, ns))
    raw = torch.rand(E, ed, generator=gen, dtype=torch.float64).requires_grad_(True)
    lin = torch.nn.Linear(ed, H * C).double()
    g = torch.randn(nd, H, C, generator=gen, dtype=torch.float64)
    q_, k_, v_ = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = og.gt_conv_unfused(q_, k_, v_, lin(raw).view(E, H, C), ei, (ns, nd))
    out.backward(g)
    got = og.gt_conv_edge_folded_f64(q, k, v, raw, lin.weight, lin.bias, ei, g=g)
    ref = {"out": out, "dq": q_.grad, "dk": k_.grad, "dv": v_.grad, "dW": lin.weight.grad, "db": lin.bias.grad, "draw": raw.grad}
    for key, want in ref.items():
        want = want.detach().numpy()
        assert np.abs(got[key] - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), key
