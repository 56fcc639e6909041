"""RawCTCNet drop-in (reference: modules/raw_ctcnet.py)."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard
from ..ops import EPI_LEAKY
from . import _stack
from .block import ResidualBlock


class RawCTCNet(nn.Module):
    """Raw 1-channel signal -> featuriser (Conv1d(1,F,fk,pad fk-1) + LeakyReLU + 1x1 + LeakyReLU; the output
    is fk-1 frames LONGER than the input, raw_ctcnet.py:57-61) -> optional position mixing -> input block +
    residual stack with skip bottlenecks -> output block -> optional softmax (raw_ctcnet.py:117-153)."""

    def __init__(self, num_features, feature_kwidth, num_labels, layers, out_dim,
                 input_kernel_size=2, input_dilation=1, positions=False, softmax=True, causal=False):
        super(RawCTCNet, self).__init__()
        self.num_features = num_features
        self.feature_kwidth = feature_kwidth
        self.num_labels = num_labels
        self.layers = layers
        self.num_layers = len(layers)
        self.out_dim = out_dim
        self.input_kernel_size = input_kernel_size
        self.input_dilation = input_dilation
        self.positions = positions
        self.softmax = softmax
        self.causal = causal

        self.feature_layer = nn.Sequential(
            nn.Conv1d(1, num_features, kernel_size=feature_kwidth, padding=(feature_kwidth - 1), dilation=1),
            nn.LeakyReLU(0.01),
            nn.Conv1d(num_features, num_features, kernel_size=1, padding=0, dilation=1),
            nn.LeakyReLU(0.01))
        if self.positions:
            self.positions_conv1x1 = nn.Sequential(
                nn.Conv1d(1, num_features, kernel_size=1, padding=0, dilation=1), nn.Hardtanh())
        self.input_block = ResidualBlock(num_features, layers[0][0], input_kernel_size, input_dilation,
                                         causal=self.causal)
        self.input_skip_bottleneck = nn.Conv1d(layers[0][0], out_dim, kernel_size=1, padding=0, dilation=1)
        blocks, necks = [], []
        for (c_in, c_out, k, d) in layers:
            blocks.append(ResidualBlock(c_in, c_out, k, d, causal=self.causal))
            necks.append(nn.Conv1d(c_out, out_dim, kernel_size=1, padding=0, dilation=1))
        self.convolutions = nn.ModuleList(blocks)
        self.bottlenecks = nn.ModuleList(necks)
        self.output_block = _stack.make_output_head(out_dim, num_labels)

        # init (raw_ctcnet.py:91-114): kaiming weights, N(0, 1e-4) biases; bottlenecks / positions = identity
        # plus 1e-4 noise.  input_skip_bottleneck is left at its default init, as in the reference.
        eps = 0.0001
        noisy = lambda p: p.data.zero_().add_(torch.randn(p.size()).mul_(eps))

        def near_identity(params):
            for p in params:
                if p.dim() > 1:
                    nn.init.eye_(p.view(p.size(0), p.size(1)))
                    p.data.add_(torch.randn(p.size()).mul_(eps))
                if p.dim() == 1:
                    noisy(p)

        if self.positions:
            near_identity(self.positions_conv1x1.parameters())
        _stack.kaiming_weights_(self.feature_layer.parameters(), noisy)
        _stack.kaiming_weights_(self.input_block.parameters(), noisy)
        _stack.kaiming_weights_(self.convolutions.parameters(), noisy)
        near_identity(self.bottlenecks.parameters())
        _stack.kaiming_weights_(self.output_block.parameters(), noisy)

    def featurize(self, seq):
        fk = self.feature_kwidth
        f0, f2 = self.feature_layer[0], self.feature_layer[2]
        t_out = seq.size(2) + fk - 1
        h = WF.conv_taps(seq, f0.weight, f0.bias, [j - (fk - 1) for j in range(fk)], epilogue=EPI_LEAKY,
                         T_out=t_out)
        return WF.conv_taps(h, f2.weight, f2.bias, [0], epilogue=EPI_LEAKY)

    @device_guard
    def forward(self, seq, t0=0):
        """seq: (batch, 1, T) -> (batch, num_labels, T + feature_kwidth - 1).  `t0` (extension, default 0)
        is the global frame index of seq[..., 0] for position mixing on a time shard."""
        from .. import fastpath
        y = fastpath.try_raw_ctcnet_forward(self, seq, t0)
        if y is not None:
            return y
        out = self.featurize(seq)
        if self.positions:
            pc = self.positions_conv1x1[0]
            out = WF.positions_mix(out, pc.weight, pc.bias, t0)
        out, skip = self.input_block(out)
        skips = WF.skip_accumulate(None, skip, self.input_skip_bottleneck.weight, self.input_skip_bottleneck.bias)
        _, skips = _stack.run_stack(out, skips, self.convolutions, self.bottlenecks)
        y = _stack.run_head(self.output_block, skips)
        return WF.channel_softmax(y) if self.softmax else y
