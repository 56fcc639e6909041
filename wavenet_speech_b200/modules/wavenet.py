"""WaveNet drop-in (reference: modules/wavenet.py)."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard
from . import _stack
from .block import ResidualBlock
from .conv_ops import CausalConv1d


class WaveNet(nn.Module):
    """Causal entry conv -> residual blocks whose skips go through per-layer 1x1 bottlenecks into a running
    sum -> LeakyReLU/1x1/LeakyReLU/1x1 -> optional channel softmax (reference wavenet.py:29-111).

    layers: list of (c_in, c_out, kernel_width, dilation)."""

    def __init__(self, in_dim, entry_kwidth, layers, out_dim, softmax=True):
        super(WaveNet, self).__init__()
        self.in_dim = in_dim
        self.entry_kwidth = entry_kwidth
        self.layers = layers
        self.num_layers = len(layers)
        self.out_dim = out_dim
        self.softmax = softmax

        self.entry_conv1d = CausalConv1d(in_dim, layers[0][0], entry_kwidth, dilation=1)
        blocks, necks = [], []
        for (c_in, c_out, kwidth, dilation) in layers:       # interleaved construction = reference RNG order
            blocks.append(ResidualBlock(c_in, c_out, kwidth, dilation))
            necks.append(nn.Conv1d(c_out, out_dim, 1, padding=0, dilation=1))
        self.convolutions = nn.ModuleList(blocks)
        self.bottlenecks = nn.ModuleList(necks)
        self.output_stack = _stack.make_output_head(out_dim, out_dim)

        zero = lambda p: p.data.zero_()
        _stack.kaiming_weights_(self.entry_conv1d.parameters(), zero)
        _stack.kaiming_weights_(self.convolutions.parameters(), zero)
        for p in self.bottlenecks.parameters():
            # the reference asks for eye-init on 2-D parameters only; bottleneck weights are 3-D, so they keep
            # nn.Conv1d's default init (wavenet.py:80-82) -- reproduced, not "fixed"
            if p.dim() == 2:
                nn.init.eye_(p)
            if p.dim() == 1:
                zero(p)
        _stack.kaiming_weights_(self.output_stack.parameters(), zero)

    @device_guard
    def forward_levels(self, levels, out_dtype=torch.bfloat16):
        """forward(one_hot(levels)) for a (B, T) integer tensor of quantised levels, without the one-hot tensor
        (tensor-core path only; not in the reference, whose loaders one-hot on the host: fns.py:6-15)."""
        from .. import fastpath
        return fastpath.wavenet_forward_levels(self, levels, out_dtype)

    @device_guard
    def forward(self, signal):
        from .. import fastpath
        y = fastpath.try_wavenet_forward(self, signal)
        if y is not None:
            return y
        out = self.entry_conv1d(signal)
        _, skips = _stack.run_stack(out, None, self.convolutions, self.bottlenecks)
        y = _stack.run_head(self.output_stack, skips)
        return WF.channel_softmax(y) if self.softmax else y
