"""WaveNetClassifier drop-in (reference: modules/classifier.py)."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard
from . import _stack
from .block import ResidualBlock


class WaveNetClassifier(nn.Module):
    """AvgPool1d(pool) -> non-causal input block + residual stack with skip bottlenecks -> output block ->
    optional softmax (reference classifier.py:17-120).  Output length is floor(T / pool)."""

    def __init__(self, in_dim, num_labels, layers, out_dim, pool_kernel_size=2, input_kernel_size=2,
                 input_dilation=1, softmax=True):
        super(WaveNetClassifier, self).__init__()
        self.in_dim = in_dim
        self.num_labels = num_labels
        self.layers = layers
        self.num_layers = len(layers)
        self.out_dim = out_dim
        self.pool_kernel_size = pool_kernel_size
        self.pool_padding = 0
        self.input_kernel_size = input_kernel_size
        self.input_dilation = input_dilation
        self.softmax = softmax

        self.mean_pool = nn.AvgPool1d(kernel_size=pool_kernel_size, padding=self.pool_padding)  # config holder
        self.input_block = ResidualBlock(in_dim, layers[0][0], input_kernel_size, input_dilation, causal=False)
        self.input_skip_bottleneck = nn.Conv1d(layers[0][0], out_dim, kernel_size=1, padding=0, dilation=1)
        blocks, necks = [], []
        for (c_in, c_out, k, d) in layers:
            blocks.append(ResidualBlock(c_in, c_out, k, d, causal=False))
            necks.append(nn.Conv1d(c_out, out_dim, kernel_size=1, padding=0, dilation=1))
        self.convolutions = nn.ModuleList(blocks)
        self.bottlenecks = nn.ModuleList(necks)
        self.output_block = _stack.make_output_head(out_dim, num_labels)

        zero = lambda p: p.data.zero_()
        _stack.kaiming_weights_(self.input_block.parameters(), zero)
        _stack.kaiming_weights_(self.convolutions.parameters(), zero)
        for p in self.bottlenecks.parameters():      # 3-D weights: eye-init never triggers (classifier.py:83-85)
            if p.dim() == 2:
                nn.init.eye_(p)
            if p.dim() == 1:
                zero(p)
        _stack.kaiming_weights_(self.output_block.parameters(), zero)

    @device_guard
    def forward(self, seq):
        from .. import fastpath
        y = fastpath.try_classifier_forward(self, seq)
        if y is not None:
            return y
        out = WF.avg_pool(seq, self.pool_kernel_size)
        out, skip = self.input_block(out)
        skips = WF.skip_accumulate(None, skip, self.input_skip_bottleneck.weight, self.input_skip_bottleneck.bias)
        _, skips = _stack.run_stack(out, skips, self.convolutions, self.bottlenecks)
        y = _stack.run_head(self.output_block, skips)
        return WF.channel_softmax(y) if self.softmax else y
