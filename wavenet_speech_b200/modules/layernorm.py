"""LayerNorm over the channel axis of a (batch, channels, time) tensor (reference: modules/layernorm.py)."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard


class LayerNorm(nn.Module):
    """gamma * (x - mean) / (std + eps) + beta with the UNBIASED std and eps added to the std, exactly as
    the reference (layernorm.py:25-28).  gamma/beta have shape (1, features, 1)."""

    def __init__(self, features, dim=1, eps=1e-6):
        super(LayerNorm, self).__init__()
        self.gamma = nn.Parameter(torch.ones(features).unsqueeze(0).unsqueeze(2))
        self.beta = nn.Parameter(torch.zeros(features).unsqueeze(0).unsqueeze(2))
        self.eps = eps
        self.dim = dim

    @device_guard
    def forward(self, x):
        if self.dim != 1 or x.dim() != 3:
            raise NotImplementedError("LayerNorm kernel normalises dim=1 of a (B, C, T) tensor")
        return WF.layer_norm(x, self.gamma, self.beta, self.eps)
