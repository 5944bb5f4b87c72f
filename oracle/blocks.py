"""TEST INFRASTRUCTURE ONLY -- functional CPU restatement of the reference's block forwards (single rank).

Follows /root/reference/src/anemoi/models/layers/block.py.  Parameters are passed as a state_dict `p`
with the reference's own key names, so the same dict loads into the reference block, into this
restatement and into the product block.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .gtconv import _ACT, graph_conv_unfused, gt_conv_unfused, mlp_forward

F = torch.nn.functional


def _ln(x, p, name):
    return F.layer_norm(x, (x.shape[-1],), p[f"{name}.weight"], p[f"{name}.bias"], 1e-5)


def _lin(x, p, name):
    return F.linear(x, p[f"{name}.weight"], p.get(f"{name}.bias"))


def _node_mlp(x, p, name, activation):
    """nn.Sequential(LayerNorm, Linear, act, Linear) (block.py:349-354)."""
    h = _ln(x, p, f"{name}.0")
    h = _ACT[activation](_lin(h, p, f"{name}.1"))
    return _lin(h, p, f"{name}.3")


def gt_mapper_block(p: dict, x, edge_attr: Tensor, edge_index: Tensor, num_heads: int, size=None,
                    activation: str = "GELU", update_src_nodes: bool = False):
    """GraphTransformerMapperBlock.forward (block.py:479-550), model_comm_group=None, num_chunks=1."""
    x_skip = x
    xs, xd = _ln(x[0], p, "layer_norm1"), _ln(x[1], p, "layer_norm2")  # block.py:491-494
    x_r = _lin(xd, p, "lin_self")
    H = num_heads
    C = p["lin_query.weight"].shape[0] // H
    q = _lin(xd, p, "lin_query").view(-1, H, C)
    k = _lin(xs, p, "lin_key").view(-1, H, C)
    v = _lin(xs, p, "lin_value").view(-1, H, C)
    e = _lin(edge_attr, p, "lin_edge").view(-1, H, C)
    out = gt_conv_unfused(q, k, v, e, edge_index, size, out_channels=C)  # block.py:526
    out = out.reshape(out.shape[0], H * C)
    out = _lin(out + x_r, p, "projection")  # block.py:531
    out = out + x_skip[1]
    dst_new = _node_mlp(out, p, "node_dst_mlp", activation) + out  # block.py:536-538
    src_new = x_skip[0]
    if update_src_nodes:
        src_new = _node_mlp(x_skip[0], p, "node_src_mlp", activation) + x_skip[0]  # block.py:540-544
    return (src_new, dst_new), edge_attr


def gt_processor_block(p: dict, x: Tensor, edge_attr: Tensor, edge_index: Tensor, num_heads: int, size=None,
                       activation: str = "GELU"):
    """GraphTransformerProcessorBlock.forward (block.py:602-635), model_comm_group=None."""
    x_skip = x
    xn = _ln(x, p, "layer_norm1")
    x_r = _lin(xn, p, "lin_self")
    H = num_heads
    C = p["lin_query.weight"].shape[0] // H
    q = _lin(xn, p, "lin_query").view(-1, H, C)
    k = _lin(xn, p, "lin_key").view(-1, H, C)
    v = _lin(xn, p, "lin_value").view(-1, H, C)
    e = _lin(edge_attr, p, "lin_edge").view(-1, H, C)
    out = gt_conv_unfused(q, k, v, e, edge_index, size, out_channels=C).reshape(-1, H * C)
    out = _lin(out + x_r, p, "projection")
    out = out + x_skip
    return _node_mlp(out, p, "node_dst_mlp", activation) + out, edge_attr


def graphconv_processor_block(p: dict, x: Tensor, edge_attr: Tensor, edge_index: Tensor, size=None,
                              mlp_extra_layers: int = 0, activation: str = "SiLU"):
    """GraphConvProcessorBlock.forward (block.py:193-223), model_comm_group=None, num_chunks=1."""
    out, edges_new = graph_conv_unfused(x, edge_attr, edge_index, p, "conv.edge_mlp.", mlp_extra_layers, activation, size)
    nodes_new = mlp_forward(torch.cat([x, out], dim=1), p, "node_mlp.", mlp_extra_layers, activation) + x
    return nodes_new, edges_new


def graphconv_mapper_block(p: dict, x, edge_attr: Tensor, edge_index: Tensor, size=None, mlp_extra_layers: int = 0,
                           activation: str = "SiLU", update_src_nodes: bool = True):
    """GraphConvMapperBlock.forward (block.py:249-286), model_comm_group=None, num_chunks=1."""
    out, edges_new = graph_conv_unfused(x, edge_attr, edge_index, p, "conv.edge_mlp.", mlp_extra_layers, activation, size)
    dst_new = mlp_forward(torch.cat([x[1], out], dim=1), p, "node_mlp.", mlp_extra_layers, activation) + x[1]
    src_new = x[0]
    if update_src_nodes:
        src_new = mlp_forward(torch.cat([x[0], x[0]], dim=1), p, "node_mlp.", mlp_extra_layers, activation) + x[0]
    return (src_new, dst_new), edges_new
