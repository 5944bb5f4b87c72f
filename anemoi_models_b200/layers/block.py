"""Drop-in graph blocks (reference layers/block.py:108-635): same class names, constructor arguments, parameter /
state_dict names and `forward` signatures, with the edge path on the fused CUDA kernels.

Single rank: LayerNorm + dense q/k/v/self/edge projections (tensor-core GEMMs through nn.Linear) -> fused conv ->
projection + node MLP, as in the reference.

Model-parallel group of P ranks: the reference shards heads (4+1 all-to-alls per GT block, every rank holding the whole
edge list) or all-gathers node features (GraphConv blocks).  Here every rank keeps its dst rows and all edges
pointing into them; only the halo of src rows is exchanged (see distributed/halo.py).  Inputs/outputs keep the
reference's sharded layouts and `shapes` bookkeeping, so mappers/processors/models call these blocks unchanged.
"""
from __future__ import annotations

import logging
import os
from abc import ABC, abstractmethod
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn

from ..distributed.collectives import shard_tensor, sync_tensor
from ..distributed.halo import (HaloPlan, build_bipartite_halo_plan, build_local_halo_plan, cached_plan, halo_exchange,
                                halo_gather, select_sharded_edges)
from ..distributed.shapes import bounds_from_shapes
from ..distributed.transformer import shard_heads, shard_sequence
from .. import gemm as tcg
from ..graph import check_edge_index, get_csr, resolve_size
from .conv import GraphConv, GraphTransformerConv
from .mlp import MLP, activation_class

LOGGER = logging.getLogger(__name__)

# Number of mapper chunks used during inference -- read at import like the reference (block.py:39).  The fused conv
# has no E x D temporaries, so chunking changes neither memory nor results; the variable stays accepted.
NUM_CHUNKS_INFERENCE = int(os.environ.get("ANEMOI_INFERENCE_NUM_CHUNKS", "1"))
_EDGE_FOLD = os.environ.get("AB2_EDGE_FOLD", "0") == "1"  # round-2 draft path, see ops.gt_conv_folded


def _autocast_once(x: Tensor) -> Tensor:
    """Under autocast every nn.Linear casts its input to the autocast dtype; LayerNorm outputs fp32.  Casting the
    normalised rows ONCE before they feed several projections gives bit-identical results and saves the repeated
    full-tensor casts (4 per block, each a pass over [N, D])."""
    if x.is_cuda and torch.is_autocast_enabled("cuda"):
        return x.to(torch.get_autocast_dtype("cuda"))
    return x


def _linear_padded_k(lin: nn.Linear, x: Tensor, multiple: int = 8) -> Tensor:
    """lin(x) with the contraction dim zero-padded to a multiple of 8 (exactly the same sums).  edge_dim is 11 in the
    reference configs; an unaligned K sends a bf16 GEMM to a legacy alignment-1 kernel (measured: 48 ms of a 330 ms
    AIFS-like step for lin_edge and its backward)."""
    k = x.shape[-1]
    pad = (-k) % multiple
    if pad == 0 or not x.is_cuda:
        return lin(x)
    F = torch.nn.functional
    return F.linear(F.pad(x, (0, pad)), F.pad(lin.weight, (0, pad)), lin.bias)


def _lin_edge(lin: nn.Linear, edge_attr: Tensor, tc: bool) -> Tensor:
    """`lin_edge(edge_attr)` (reference block.py:497, 618).  edge_dim is 11 in the reference configs: cuBLASLt runs the forward
    and both backward GEMMs of such a skinny contraction far below the HBM rate (measured on the headline graph: 4.9 ms per
    forward + backward for 1.5 GB of output, bench.py --workload edgepath); on the tcgen05 kernel they are three bandwidth-bound
    launches (the contraction zero-padded to 16 so that rows are 16-byte aligned; TMA fills the rest of the 64-wide k-block)."""
    if not tc or edge_attr.dim() != 2:
        return _linear_padded_k(lin, edge_attr)
    F = torch.nn.functional
    pad = (-edge_attr.shape[-1]) % 16
    x = tcg.pad_cast(edge_attr, 16) if (pad or edge_attr.dtype != torch.bfloat16) else edge_attr  # one kernel: pad + bf16
    w = F.pad(lin.weight, (0, pad)) if pad else lin.weight
    return tcg.linear_wb(x, w, lin.bias)


def _fold_applies(query: Tensor, edge_attr: Tensor) -> bool:
    """round-2 switch AB2_EDGE_FOLD=1: lin_edge folded into the conv (ops.gt_conv_folded) when the raw features fit its 16 columns"""
    return _EDGE_FOLD and query.is_cuda and edge_attr.dim() == 2 and edge_attr.shape[1] < 16


def _group_active(group) -> bool:
    return group is not None and bool(group) and dist.get_world_size(group=group) > 1


class BaseBlock(nn.Module, ABC):
    """Base class for network blocks (reference block.py:42-58)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    @abstractmethod
    def forward(self, x, edge_attr, edge_index, shapes, batch_size, size=None, model_comm_group=None): ...


# ------------------------------------------------------------------------------------------------------------
# GraphConv blocks
# ------------------------------------------------------------------------------------------------------------
class GraphConvBaseBlock(BaseBlock):
    """Message passing block with MLPs for node embeddings (reference block.py:108-167)."""

    def __init__(self, in_channels: int, out_channels: int, mlp_extra_layers: int = 0, activation: str = "SiLU",
                 update_src_nodes: bool = True, num_chunks: int = 1, **kwargs) -> None:
        super().__init__(**kwargs)
        self.update_src_nodes = update_src_nodes
        self.num_chunks = num_chunks
        self.node_mlp = MLP(2 * in_channels, out_channels, out_channels, n_extra_layers=mlp_extra_layers, activation=activation)
        self.conv = GraphConv(in_channels=in_channels, out_channels=out_channels, mlp_extra_layers=mlp_extra_layers,
                              activation=activation)

    def _conv(self, x_src: Tensor, x_dst: Tensor, edge_attr: Tensor, edge_index: Tensor, shapes_src, shapes_dst,
              model_comm_group, size):
        """out rows of THIS rank's dst shard and the updated edge features of this rank's edges.
        (The reference's `num_chunks` loop over edge slices, block.py:205-215, only bounds the size of its [E,3D]
        temporaries; the fused path has none, so one call gives the same sums.)"""
        if not _group_active(model_comm_group):
            single = x_src is x_dst
            return self.conv(x_src if single else (x_src, x_dst), edge_attr, edge_index, size=size)
        # sharded: x_src / x_dst are this rank's rows; edge_index holds this rank's 1-hop edges with GLOBAL ids
        check_edge_index(edge_index)
        sb, db = bounds_from_shapes(shapes_src), bounds_from_shapes(shapes_dst)
        rank = dist.get_rank(group=model_comm_group)
        plan: HaloPlan = cached_plan(self, ("local", tuple(sb), tuple(db), rank), edge_index, model_comm_group,
                                     lambda: build_local_halo_plan(edge_index, sb, db, model_comm_group))
        x_need = halo_gather(x_src, plan, model_comm_group)  # replaces sync_tensor's full all-gather (block.py:203)
        return self.conv((x_need, x_dst), edge_attr, plan.local_edge_index, size=(plan.n_src, plan.num_dst_local))

    def _node_update(self, x_in: Tensor, x_res: Tensor) -> Tensor:
        """`node_mlp(x_in) + x_res` (reference block.py:217-219, 277-283): Linear-act-Linear-act-Linear-LayerNorm; bf16 CUDA inputs
        run on the tcgen05 GEMMs (activations in the epilogues) and the LayerNorm kernel, anything else on the nn modules."""
        body = self.node_mlp.model
        layers = list(body) if isinstance(body, nn.Sequential) else None
        if (layers is not None and isinstance(layers[-1], nn.LayerNorm) and layers[-1].elementwise_affine
                and type(layers[1]).__name__ in tcg.ACT_CODES
                and tcg.tc_applies(x_in, *[d for m in layers if isinstance(m, nn.Linear) for d in (m.in_features, m.out_features)])):
            linears = [m for m in layers if isinstance(m, nn.Linear)]
            y = tcg.mlp_forward(x_in, linears, tcg.ACT_CODES[type(layers[1]).__name__])
            return tcg.layer_norm(y, layers[-1], out_dtype=y.dtype) + x_res
        return self.node_mlp(x_in) + x_res

    @abstractmethod
    def forward(self, x, edge_attr, edge_index, shapes, model_comm_group=None, size=None): ...


class GraphConvProcessorBlock(GraphConvBaseBlock):
    """reference block.py:170-223"""

    def forward(self, x: Tensor, edge_attr: Tensor, edge_index: Tensor, shapes: tuple, model_comm_group=None,
                size: Optional[Tuple[int, int]] = None):
        out, edges_new = self._conv(x, x, edge_attr, edge_index, shapes[1], shapes[1], model_comm_group, size)
        nodes_new = self._node_update(torch.cat([x, out], dim=1), x)
        return nodes_new, edges_new


class GraphConvMapperBlock(GraphConvBaseBlock):
    """reference block.py:226-286"""

    def forward(self, x: Tuple[Tensor, Tensor], edge_attr: Tensor, edge_index: Tensor, shapes: tuple,
                model_comm_group=None, size: Optional[Tuple[int, int]] = None):
        out, edges_new = self._conv(x[0], x[1], edge_attr, edge_index, shapes[0], shapes[1], model_comm_group, size)
        nodes_new_dst = self._node_update(torch.cat([x[1], out], dim=1), x[1])
        # update only needed in forward mapper
        nodes_new_src = x[0] if not self.update_src_nodes else self._node_update(torch.cat([x[0], x[0]], dim=1), x[0])
        return (nodes_new_src, nodes_new_dst), edges_new


# ------------------------------------------------------------------------------------------------------------
# GraphTransformer blocks
# ------------------------------------------------------------------------------------------------------------
class GraphTransformerBaseBlock(BaseBlock, ABC):
    """Message passing block with MLPs for node embeddings (reference block.py:289-426)."""

    def __init__(self, in_channels: int, hidden_dim: int, out_channels: int, edge_dim: int, num_heads: int = 16,
                 bias: bool = True, activation: str = "GELU", num_chunks: int = 1, update_src_nodes: bool = False,
                 **kwargs) -> None:
        super().__init__(**kwargs)
        self.update_src_nodes = update_src_nodes
        self.out_channels_conv = out_channels // num_heads
        self.num_heads = num_heads
        self.num_chunks = num_chunks

        self.lin_key = nn.Linear(in_channels, num_heads * self.out_channels_conv)
        self.lin_query = nn.Linear(in_channels, num_heads * self.out_channels_conv)
        self.lin_value = nn.Linear(in_channels, num_heads * self.out_channels_conv)
        self.lin_self = nn.Linear(in_channels, num_heads * self.out_channels_conv, bias=bias)
        self.lin_edge = nn.Linear(edge_dim, num_heads * self.out_channels_conv)

        self.conv = GraphTransformerConv(out_channels=self.out_channels_conv)
        self.projection = nn.Linear(out_channels, out_channels)

        act_func = activation_class(activation)
        self.node_dst_mlp = nn.Sequential(nn.LayerNorm(out_channels), nn.Linear(out_channels, hidden_dim), act_func(),
                                          nn.Linear(hidden_dim, out_channels))
        self.layer_norm1 = nn.LayerNorm(in_channels)
        if self.update_src_nodes:
            self.node_src_mlp = nn.Sequential(nn.LayerNorm(out_channels), nn.Linear(out_channels, hidden_dim), act_func(),
                                              nn.Linear(hidden_dim, out_channels))

    # -- dense node-side contractions: tcgen05 kernels (gemm.py / csrc/gemm_tc.cu) for CUDA bf16 (or bf16-autocast) inputs,
    #    nn.LayerNorm / nn.Linear otherwise (fp32 parity runs, CPU tests with the oracle conv)
    def _tc(self, x: Tensor) -> bool:
        lin2 = self.node_dst_mlp[3]
        return (tcg.tc_applies(x, self.lin_key.in_features, self.lin_key.out_features, lin2.in_features, lin2.out_features)
                and type(self.node_dst_mlp[2]).__name__ in tcg.ACT_CODES and self.layer_norm1.elementwise_affine)

    def _ln(self, ln: nn.LayerNorm, x: Tensor, tc: bool) -> Tensor:
        return tcg.layer_norm(x, ln) if tc else _autocast_once(ln(x))

    def _ln_fork(self, ln: nn.LayerNorm, x: Tensor, tc: bool) -> Tuple[Tensor, Tensor]:
        """(ln(x), x as the skip connection): on the tensor-core path the skip gradient is added inside the LayerNorm backward"""
        return tcg.layer_norm_fork(x, ln) if tc else (_autocast_once(ln(x)), x)

    def _lin_multi(self, lins, x: Tensor, tc: bool):
        """[lin(x) for lin in lins]: one GEMM forward, one dgrad, one wgrad on the tensor-core path (gemm.linear_multi)"""
        return tcg.linear_multi(x, lins) if tc else [lin(x) for lin in lins]

    def _lin(self, lin: nn.Linear, x: Tensor, tc: bool, residual: Optional[Tensor] = None) -> Tensor:
        if tc:
            return tcg.linear(x, lin, residual=residual)
        y = lin(x)
        return y if residual is None else y + residual

    def _mlp_res(self, mlp: nn.Sequential, x: Tensor, tc: bool) -> Tensor:
        """`mlp(x) + x` for mlp = Sequential(LayerNorm, Linear, act, Linear) (reference block.py:349-354, 537, 633)."""
        if not tc:
            return mlp(x) + x
        act = tcg.ACT_CODES[type(mlp[2]).__name__]
        xn, x_skip = tcg.layer_norm_fork(x, mlp[0])
        pre, h = tcg.linear(xn, mlp[1], act_out=act)
        return tcg.act_linear(pre, h, mlp[3], act, residual=x_skip)

    # -- API compatibility (reference block.py:366-414).  forward() does not use the head all-to-all with a group (it shards
    #    by dst rows and exchanges a halo, see _attend); on one rank these are pure reshapes.
    def shard_qkve_heads(self, query, key, value, edges, shapes, batch_size, model_comm_group=None):
        """[(batch grid), heads*vars] rows of a sequence shard -> [(batch grid_total), heads_of_this_rank, vars]."""
        H, C = self.num_heads, self.out_channels_conv
        if not _group_active(model_comm_group):
            return tuple(t.reshape(t.shape[0], H, C) for t in (query, key, value, edges))
        shape_src, shape_dst, shape_edges = shapes
        out = []
        for t, shp in ((query, shape_dst), (key, shape_src), (value, shape_src), (edges, shape_edges)):
            t = t.reshape(batch_size, -1, H, C).transpose(1, 2)  # batch heads grid vars
            t = shard_heads(t, shp, model_comm_group)
            out.append(t.transpose(1, 2).reshape(-1, t.shape[1], C))
        return tuple(out)

    def shard_output_seq(self, out, shapes, batch_size, model_comm_group=None):
        """[(batch grid_total), heads_of_this_rank, vars] -> [(batch grid_shard), heads*vars]."""
        if not _group_active(model_comm_group):
            return out.reshape(out.shape[0], -1)
        t = out.reshape(batch_size, -1, out.shape[-2], out.shape[-1]).transpose(1, 2)  # batch heads grid vars
        t = shard_sequence(t, shapes[1], model_comm_group)
        return t.transpose(1, 2).reshape(-1, t.shape[1] * t.shape[-1])

    def _attend(self, query: Tensor, key: Tensor, value: Tensor, edge_attr: Tensor, edge_index: Tensor, shapes: tuple,
                batch_size: int, model_comm_group, size) -> Tensor:
        """conv output rows [(own dst rows), H*C] for projected q [Nd_r,D], k/v [Ns_r,D] and RAW edge_attr."""
        H, C = self.num_heads, self.out_channels_conv
        if not _group_active(model_comm_group):
            if _fold_applies(query, edge_attr):
                # round-2 draft (AB2_EDGE_FOLD=1, off by default): lin_edge folded into the conv, no [E, H*C] edge tensor
                from .. import ops

                check_edge_index(edge_index)
                ns, nd = resolve_size(size, key.shape[0], query.shape[0])
                out = ops.gt_conv_folded(query.view(-1, H, C), key.view(-1, H, C), value.view(-1, H, C), edge_attr,
                                         self.lin_edge.weight, self.lin_edge.bias, get_csr(edge_index, ns, nd))
                return out.reshape(out.shape[0], H * C)
            edges = _lin_edge(self.lin_edge, edge_attr, query.dtype == torch.bfloat16 and self._tc(query))
            q, k, v, e = self.shard_qkve_heads(query, key, value, edges, shapes, batch_size, model_comm_group)
            out = self.conv(query=q, key=k, value=v, edge_attr=e, edge_index=edge_index, size=size)
            return self.shard_output_seq(out, shapes, batch_size, model_comm_group)
        assert batch_size == 1, "Only batch size of 1 is supported when model is sharded across GPUs"
        check_edge_index(edge_index)
        shapes_src, shapes_dst, shapes_edge = shapes
        sb, db = bounds_from_shapes(shapes_src), bounds_from_shapes(shapes_dst)
        rank = dist.get_rank(group=model_comm_group)
        plan: HaloPlan = cached_plan(self, ("bipartite", tuple(sb), tuple(db), rank), edge_index, model_comm_group,
                                     lambda: build_bipartite_halo_plan(edge_index, sb, db, rank))
        # raw attributes of the edges this rank owns (they arrive sharded by original edge order), then project locally
        ea_local = select_sharded_edges(edge_attr, shapes_edge, plan.edge_ids, model_comm_group, plan)
        e = _lin_edge(self.lin_edge, ea_local, query.dtype == torch.bfloat16 and self._tc(query)).view(-1, H, C)
        q, k, v = query.view(-1, H, C), key.view(-1, H, C), value.view(-1, H, C)
        size = (plan.n_src, plan.num_dst_local)
        if q.is_cuda:
            from .. import ops

            csr = get_csr(plan.local_edge_index, *size)
            out = ops.gt_conv_sharded(q, k, v, e, csr, plan, model_comm_group)
        else:  # generic composition (differentiable exchange + conv on [own | halo] rows)
            halo = (halo_exchange(k, plan, model_comm_group), halo_exchange(v, plan, model_comm_group))
            out = self.conv(query=q, key=k, value=v, edge_attr=e, edge_index=plan.local_edge_index, size=size, halo=halo)
        return out.reshape(out.shape[0], H * C)

    @abstractmethod
    def forward(self, x, edge_attr, edge_index, shapes, batch_size, model_comm_group=None, size=None): ...


class GraphTransformerMapperBlock(GraphTransformerBaseBlock):
    """Graph Transformer Block for node embeddings, bipartite (reference block.py:429-550)."""

    def __init__(self, in_channels: int, hidden_dim: int, out_channels: int, edge_dim: int, num_heads: int = 16,
                 bias: bool = True, activation: str = "GELU", num_chunks: int = 1, update_src_nodes: bool = False,
                 **kwargs) -> None:
        super().__init__(in_channels=in_channels, hidden_dim=hidden_dim, out_channels=out_channels, edge_dim=edge_dim,
                         num_heads=num_heads, bias=bias, activation=activation, num_chunks=num_chunks,
                         update_src_nodes=update_src_nodes, **kwargs)
        self.layer_norm2 = nn.LayerNorm(in_channels)

    def forward(self, x: Tuple[Tensor, Tensor], edge_attr: Tensor, edge_index: Tensor, shapes: tuple, batch_size: int,
                model_comm_group=None, size: Optional[Tuple[int, int]] = None):
        tc = self._tc(x[0])
        (x_src, skip_src), (x_dst, skip_dst) = self._ln_fork(self.layer_norm1, x[0], tc), self._ln_fork(self.layer_norm2, x[1], tc)
        x_skip = (skip_src, skip_dst)
        query, x_r = self._lin_multi((self.lin_query, self.lin_self), x_dst, tc)
        key, value = self._lin_multi((self.lin_key, self.lin_value), x_src, tc)

        if model_comm_group is not None:
            assert (model_comm_group.size() == 1 or batch_size == 1), \
                "Only batch size of 1 is supported when model is sharded across GPUs"

        out = self._attend(query, key, value, edge_attr, edge_index, shapes, batch_size, model_comm_group, size)

        out = self._lin(self.projection, out + x_r, tc, residual=x_skip[1])  # projection(out + x_r) + x_skip: one epilogue
        nodes_new_dst = self._mlp_res(self.node_dst_mlp, out, tc)
        nodes_new_src = self._mlp_res(self.node_src_mlp, x_skip[0], tc) if self.update_src_nodes else x_skip[0]
        return (nodes_new_src, nodes_new_dst), edge_attr


class GraphTransformerProcessorBlock(GraphTransformerBaseBlock):
    """Graph Transformer Block for node embeddings, one node set (reference block.py:553-635)."""

    def forward(self, x: Tensor, edge_attr: Tensor, edge_index: Tensor, shapes: tuple, batch_size: int,
                model_comm_group=None, size: Optional[Tuple[int, int]] = None):
        tc = self._tc(x)
        x, x_skip = self._ln_fork(self.layer_norm1, x, tc)
        query, key, value, x_r = self._lin_multi((self.lin_query, self.lin_key, self.lin_value, self.lin_self), x, tc)

        if model_comm_group is not None:
            assert (model_comm_group.size() == 1 or batch_size == 1), \
                "Only batch size of 1 is supported when model is sharded across GPUs"

        out = self._attend(query, key, value, edge_attr, edge_index, shapes, batch_size, model_comm_group, size)
        out = self._lin(self.projection, out + x_r, tc, residual=x_skip)
        nodes_new = self._mlp_res(self.node_dst_mlp, out, tc)
        return nodes_new, edge_attr
