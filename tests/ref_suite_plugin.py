"""pytest plugin (tests only): run the REFERENCE's own graph-path tests against this repo's blocks.

Loaded with `-p ref_suite_plugin` when pytest is pointed at /root/reference/tests: before the reference's test modules are
imported it calls `install()` (so `from anemoi.models.layers.block import GraphTransformerMapperBlock` etc. resolve to this
repo's classes) and, because these tests run on CPU tensors, puts the oracle's CPU restatement behind the conv entry points."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    for p in (ROOT, os.path.join(ROOT, "oracle", "pyg_shim"), os.path.join(ROOT, "oracle", "ref_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch

    import anemoi_models_b200 as b2
    import anemoi_models_b200.layers.conv as convmod
    from anemoi_models_b200 import ops
    from oracle import gtconv as og

    b2.install(edge_partition=False)

    class CpuPlan:
        def __init__(self, edge_index, ns, nd):
            self.edge_index, self.num_src, self.num_dst, self.num_edges = edge_index, ns, nd, edge_index.shape[1]

    def cpu_conv(q, k, v, e, plan, halo=None):
        if halo is not None:
            k, v = torch.cat([k, halo[0]]), torch.cat([v, halo[1]])
        return og.gt_conv_unfused(q, k, v, e, plan.edge_index, (plan.num_src, plan.num_dst))

    def graphconv_forward(self, x, edge_attr, edge_index, size=None, plan=None):
        p = dict(self.named_parameters())
        return og.graph_conv_unfused(x, edge_attr, edge_index, {"edge_mlp." + k[len("edge_mlp."):]: v for k, v in p.items()},
                                     "edge_mlp.", size=size)

    convmod.get_csr = lambda ei, ns, nd: CpuPlan(ei, ns, nd)
    ops.gt_conv = cpu_conv
    convmod.GraphConv.forward = graphconv_forward
    config._ab2_installed = True


def pytest_report_header(config):
    return "anemoi_models_b200 installed over the reference (blocks / convs rebound); conv arithmetic = oracle on the CPU"


_seen = {}


def pytest_collection_finish(session):
    """Which classes did the reference's test modules bind at import?  (One of them later calls importlib.reload on
    anemoi.models.layers.block -- test_block_graphtransformer.py:365 -- which restores the reference's names in THAT module's
    namespace only; the test modules, mapper.py and chunk.py keep the classes they imported.)"""
    names = ("GraphTransformerMapperBlock", "GraphTransformerProcessorBlock", "GraphConvMapperBlock", "GraphConvProcessorBlock",
             "GraphTransformerConv", "GraphConv")
    for item in session.items:
        g = vars(item.module)
        for n in names:
            if n in g:
                _seen.setdefault(n, set()).add(g[n].__module__)
    import anemoi.models.layers.chunk as ref_chunk
    import anemoi.models.layers.mapper as ref_mapper

    for mod in (ref_chunk, ref_mapper):
        for n in names:
            if hasattr(mod, n):
                _seen.setdefault(n, set()).add(getattr(mod, n).__module__)


def pytest_terminal_summary(terminalreporter):
    ok = bool(_seen) and all(all(m.startswith("anemoi_models_b200") for m in mods) for mods in _seen.values())
    terminalreporter.write_line("ab2-plugin: classes bound by the reference's tests / mapper.py / chunk.py: "
                                + ", ".join(f"{n}->{'|'.join(sorted(m))}" for n, m in sorted(_seen.items())))
    terminalreporter.write_line("ab2-plugin: ALL-FROM-ANEMOI_MODELS_B200" if ok else "ab2-plugin: SOME-FROM-REFERENCE")
