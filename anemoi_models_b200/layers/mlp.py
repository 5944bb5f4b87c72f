"""MLP with the reference's structure and state_dict keys (reference layers/mlp.py:22-89, layers/utils.py:27-39).

Dense node/edge-side contractions stay `nn.Linear` (cuBLASLt tensor-core GEMMs); the conv modules fuse what
surrounds them.
"""
from __future__ import annotations

import logging

import torch
from torch import Tensor, nn
from torch.utils.checkpoint import checkpoint

LOGGER = logging.getLogger(__name__)


class CheckpointWrapper(nn.Module):
    """reference layers/utils.py:16-24"""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return checkpoint(self.module, *args, **kwargs, use_reentrant=False)


class AutocastLayerNorm(nn.LayerNorm):
    """LayerNorm whose output is cast back to the input dtype (reference layers/utils.py:27-39)."""

    def forward(self, x: Tensor) -> Tensor:
        return super().forward(x).type_as(x)


def activation_class(activation: str):
    try:
        return getattr(nn, activation)
    except AttributeError as ae:
        LOGGER.error("Activation function %s not supported", activation)
        raise RuntimeError from ae


class MLP(nn.Module):
    """Linear, act, (Linear, act) x (n_extra_layers + 1), Linear, [act], [AutocastLayerNorm] -- reference mlp.py:74-84."""

    def __init__(self, in_features: int, hidden_dim: int, out_features: int, n_extra_layers: int = 0,
                 activation: str = "SiLU", final_activation: bool = False, layer_norm: bool = True,
                 checkpoints: bool = False) -> None:
        super().__init__()
        act = activation_class(activation)
        widths = [in_features] + [hidden_dim] * (n_extra_layers + 2)  # input layer + (n_extra_layers + 1) hidden layers
        layers = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(fan_in, fan_out), act()]
        layers.append(nn.Linear(hidden_dim, out_features))
        if final_activation:
            layers.append(act())
        if layer_norm:
            layers.append(AutocastLayerNorm(out_features))
        body = nn.Sequential(*layers)  # indices = the reference's state_dict keys (model.0, model.2, ...)
        self.model = CheckpointWrapper(body) if checkpoints else body

    def forward(self, x: Tensor) -> Tensor:
        return self.model(x)
