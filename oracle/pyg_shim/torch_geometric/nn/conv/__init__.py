"""Restated torch_geometric.nn.conv.MessagePassing (the subset the reference uses: conv.py:27,79)."""
import inspect
from typing import Optional

import torch
from torch import Tensor

from torch_geometric.utils import scatter

_SPECIAL = {"edge_index", "adj_t", "edge_index_i", "edge_index_j", "size", "size_i", "size_j", "ptr", "index", "dim_size"}


def _params(fn, skip_first=0):
    names = list(inspect.signature(fn).parameters.keys())
    return [n for n in names[skip_first:] if n not in ("self", "kwargs", "args")]


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr: Optional[str] = "add", *, flow: str = "source_to_target", node_dim: int = -2, **kwargs):
        super().__init__()
        if kwargs.pop("aggr_kwargs", None):
            raise NotImplementedError
        kwargs.pop("decomposed_layers", None)
        if kwargs:
            raise TypeError(f"unexpected kwargs {list(kwargs)}")
        if aggr not in ("add", "sum", "mean", "max", None):
            raise NotImplementedError(aggr)
        self.aggr = aggr
        self.flow = flow
        if flow not in ("source_to_target", "target_to_source"):
            raise ValueError(flow)
        self.node_dim = node_dim

    # -- helpers -----------------------------------------------------------------------------
    def _check_input(self, edge_index, size):
        if not isinstance(edge_index, Tensor):
            raise ValueError("`MessagePassing.propagate` only supports integer tensors of shape `[2, num_messages]`")
        int_dtypes = (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64)
        if edge_index.dtype not in int_dtypes:
            raise ValueError(f"Expected 'edge_index' to be of integer type (got '{edge_index.dtype}')")
        if edge_index.dim() != 2:
            raise ValueError(f"Expected 'edge_index' to be two-dimensional (got {edge_index.dim()} dimensions)")
        if edge_index.size(0) != 2:
            raise ValueError(f"Expected 'edge_index' to have size '2' in the first dimension (got '{edge_index.size(0)}')")
        the_size = [None, None]
        if size is not None:
            the_size[0], the_size[1] = size[0], size[1]
        return the_size

    def _set_size(self, size, dim, src):
        the_size = size[dim]
        if the_size is None:
            size[dim] = src.size(self.node_dim)
        elif the_size != src.size(self.node_dim):
            raise ValueError(
                f"Encountered tensor with size {src.size(self.node_dim)} in dimension {self.node_dim}, "
                f"but expected size {the_size}."
            )

    def _lift(self, src, edge_index, dim):
        return src.index_select(self.node_dim, edge_index[dim])

    def _collect(self, args, edge_index, size, kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        out = {}
        for arg in args:
            if arg[-2:] not in ("_i", "_j"):
                out[arg] = kwargs.get(arg, inspect.Parameter.empty)
            else:
                dim = j if arg[-2:] == "_j" else i
                data = kwargs.get(arg[:-2], inspect.Parameter.empty)
                if isinstance(data, (tuple, list)):
                    assert len(data) == 2
                    if isinstance(data[1 - dim], Tensor):
                        self._set_size(size, 1 - dim, data[1 - dim])
                    data = data[dim]
                if isinstance(data, Tensor):
                    self._set_size(size, dim, data)
                    data = self._lift(data, edge_index, dim)
                out[arg] = data
        out["adj_t"] = None
        out["edge_index"] = edge_index
        out["edge_index_i"] = edge_index[i]
        out["edge_index_j"] = edge_index[j]
        out["ptr"] = None
        out["index"] = out["edge_index_i"]
        out["size"] = size
        out["size_i"] = size[i] if size[i] is not None else size[j]
        out["size_j"] = size[j] if size[j] is not None else size[i]
        out["dim_size"] = out["size_i"]
        return out

    @staticmethod
    def _distribute(fn, coll, skip_first=0):
        sig = inspect.signature(fn)
        out = {}
        for idx, (name, p) in enumerate(sig.parameters.items()):
            if idx < skip_first or name in ("self",) or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
                continue
            v = coll.get(name, inspect.Parameter.empty)
            if v is inspect.Parameter.empty:
                if p.default is inspect.Parameter.empty:
                    raise TypeError(f"Required parameter {name} is empty.")
                v = p.default
            out[name] = v
        return out

    # -- API ---------------------------------------------------------------------------------
    def propagate(self, edge_index, size=None, **kwargs):
        size = self._check_input(edge_index, size)
        user_args = set(_params(self.message)) | set(_params(self.aggregate, 1)) | set(_params(self.update, 1))
        coll = self._collect(user_args - _SPECIAL, edge_index, size, kwargs)
        msg_kwargs = self._distribute(self.message, coll)
        out = self.message(**msg_kwargs)
        aggr_kwargs = self._distribute(self.aggregate, coll, skip_first=1)
        out = self.aggregate(out, **aggr_kwargs)
        upd_kwargs = self._distribute(self.update, coll, skip_first=1)
        return self.update(out, **upd_kwargs)

    def message(self, x_j: Tensor) -> Tensor:
        return x_j

    def aggregate(self, inputs: Tensor, index: Tensor, ptr: Optional[Tensor] = None, dim_size: Optional[int] = None) -> Tensor:
        reduce = "sum" if self.aggr in ("add", "sum") else self.aggr
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=reduce)

    def update(self, inputs):
        return inputs
