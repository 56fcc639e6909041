"""Residual / gated blocks (reference: modules/block.py)."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard
from ..ops import EPI_NONE
from .conv_ops import CausalConv1d, NonCausalConv1d
from .layernorm import LayerNorm


class GatedActivationUnit(nn.Module):
    """tanh(x) * sigmoid(y) (reference block.py:177-188).  Inside ResidualBlock the gate is fused into the
    contraction epilogue; called on its own it runs the stand-alone gate kernel."""

    @device_guard
    def forward(self, x, y):
        return WF.gated_activation(x, y)

    def __repr__(self):
        return self.__class__.__name__ + ' ()'


class ResidualBlock(nn.Module):
    """Two dilated convs -> tanh*sigmoid gate -> 1x1 residual (+ a learned Linear projection of the block
    input, not an identity) and 1x1 skip (reference block.py:15-82).  forward returns (residual, skip)."""

    def __init__(self, in_channels, out_channels, kernel_width, dilation, causal=True, conditioning=None):
        super(ResidualBlock, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_width = kernel_width
        self.dilation = dilation
        self.causal = causal
        self.conditioning = conditioning is not None      # accepted and ignored, as in the reference
        conv_cls = CausalConv1d if causal else NonCausalConv1d
        # construction order fixes the RNG stream -> same default init as the reference under a seed
        self.conv_tanh = conv_cls(in_channels, out_channels, kernel_width, dilation=dilation)
        self.conv_sigmoid = conv_cls(in_channels, out_channels, kernel_width, dilation=dilation)
        self.conv1x1_residual = nn.Conv1d(out_channels, out_channels, kernel_size=1)
        self.conv1x1_skip = nn.Conv1d(out_channels, out_channels, kernel_size=1)
        self.gated_activation = GatedActivationUnit()
        self.residual_proj = nn.Linear(in_channels, out_channels)
        self.receptive_field = self.conv_tanh.receptive_field

    @property
    def offsets(self):
        return self.conv_tanh.offsets

    @device_guard
    def forward(self, seq):
        return WF.residual_block(seq, self, self.offsets)


class MultiplicativeUnit(nn.Module):
    """g1 * tanh(g2*h + g3*tanh(u)) with four causal convs (reference block.py:192-225)."""

    def __init__(self, ndim, k, dilation=1):
        super(MultiplicativeUnit, self).__init__()
        self.ndim = ndim
        self.gate1 = CausalConv1d(ndim, ndim, kernel_width=k, dilation=dilation)
        self.gate2 = CausalConv1d(ndim, ndim, kernel_width=k, dilation=dilation)
        self.gate3 = CausalConv1d(ndim, ndim, kernel_width=k, dilation=dilation)
        self.update = CausalConv1d(ndim, ndim, kernel_width=k, dilation=dilation)
        self.init()
        self.receptive_field = max(self.gate1.receptive_field, self.gate2.receptive_field,
                                   self.gate3.receptive_field, self.update.receptive_field)

    @property
    def convs(self):
        return (self.gate1.conv1d, self.gate2.conv1d, self.gate3.conv1d, self.update.conv1d)

    @device_guard
    def forward(self, h):
        # inference: ONE launch (the gate is the contraction's epilogue); with gradients: contraction + gate kernel
        return WF.multiplicative_unit(h, self.convs, self.gate1.offsets)

    def init(self):
        for p in self.parameters():
            if p.dim() >= 2:
                nn.init.kaiming_normal_(p)
            if p.dim() == 1:
                p.data.zero_().add_(0.001 * torch.randn(p.size()))


class _ByteNetBlock(nn.Module):
    """`seq + stack(seq)` (block.py:117-119, 164-166).  `stack` is the reference's nn.Sequential (same indices, same
    state_dict keys); forward walks it in fused steps instead of calling it module by module."""

    def init(self):
        for p in self.parameters():
            if p.dim() >= 2:
                nn.init.kaiming_normal_(p)
            if p.dim() == 1:
                p.data.zero_().add_(0.001 * torch.randn(p.size()))


class _Conv1x1(nn.Conv1d):
    """nn.Conv1d(k=1) parameter container whose forward is the libwnb200 contraction."""

    @device_guard
    def forward(self, x):
        return WF.conv_taps(x, self.weight, self.bias, [0])


class _ReLU(nn.Module):
    @device_guard
    def forward(self, x):
        return torch.relu(x)


class ResidualMUBlock(_ByteNetBlock):
    """ByteNet residual multiplicative block (reference block.py:86-126)."""

    def __init__(self, nchannels, k_width, dilation=1):
        super(ResidualMUBlock, self).__init__()
        self.nchannels, self.k_width, self.dilation = nchannels, k_width, dilation
        half = int(nchannels / 2)
        self.stack = nn.Sequential(
            LayerNorm(nchannels), _ReLU(), _Conv1x1(nchannels, half, 1), LayerNorm(half, dim=1), _ReLU(),
            MultiplicativeUnit(half, k_width, dilation=dilation), MultiplicativeUnit(half, 1, dilation=1),
            _Conv1x1(half, nchannels, 1))
        self.receptive_field = self.stack[5].receptive_field

    @device_guard
    def forward(self, seq):
        """Three (statistics, contraction) pairs: each LayerNorm + ReLU is applied as the following convolution loads
        its operand, the last contraction's store adds `seq`."""
        from .. import bytenet_tc
        if bytenet_tc.eligible(self, seq):        # bf16 inference: the tcgen05 form (NLC, dense2 contractions)
            return bytenet_tc.relu_block_forward(self, seq)
        st = self.stack
        seq = seq.contiguous()
        h = WF.fused_conv(seq, st[2].weight, st[2].bias, [0], ln=st[0])
        h = WF.fused_conv(h, st[5].conv1d.weight, st[5].conv1d.bias, st[5].offsets, ln=st[3])
        return WF.fused_conv(h, st[8].weight, st[8].bias, [0], ln=st[6], residual=seq)

    @device_guard
    def forward(self, seq):
        """Inference, 6 launches and no normalised tensor in HBM: statistics; 1x1 reading ReLU(LN(seq)) on the fly;
        statistics; MU(k, d) = ONE contraction whose operand and whose gate's h are ReLU(LN(.)) formed on the fly and
        whose epilogue is the gate; MU(1) likewise; 1x1 whose store adds `seq`.  With gradients the two units run as
        contraction + gate kernel on a stored h (their backward needs both)."""
        from .. import bytenet_tc
        if bytenet_tc.eligible(self, seq):        # bf16 inference: the tcgen05 form (NLC, dense2 contractions)
            return bytenet_tc.mu_block_forward(self, seq)
        st = self.stack
        seq = seq.contiguous()
        z = WF.fused_conv(seq, st[2].weight, st[2].bias, [0], ln=st[0])
        mu1, mu2 = st[5], st[6]
        if WF.needs_grad(seq, *self.parameters()) or len(mu1.gate1.offsets) > WF.ops.MAX_SRC:
            h = mu2(mu1(WF.ln_relu(z, st[3])))
        else:
            h = WF.multiplicative_unit_fused(z, mu1.convs, mu1.gate1.offsets, ln=st[3])
            h = WF.multiplicative_unit_fused(h, mu2.convs, mu2.gate1.offsets)
        return WF.fused_conv(h, st[7].weight, st[7].bias, [0], residual=seq)


class ResidualReLUBlock(_ByteNetBlock):
    """ByteNet residual ReLU block (reference block.py:130-173)."""

    def __init__(self, nchannels, k_width, dilation=1):
        super(ResidualReLUBlock, self).__init__()
        self.nchannels, self.k_width, self.dilation = nchannels, k_width, dilation
        half = int(nchannels / 2)
        self.stack = nn.Sequential(
            LayerNorm(nchannels), _ReLU(), _Conv1x1(nchannels, half, 1), LayerNorm(half), _ReLU(),
            CausalConv1d(half, half, kernel_width=k_width, dilation=dilation), LayerNorm(half), _ReLU(),
            _Conv1x1(half, nchannels, 1))
        self.receptive_field = self.stack[5].receptive_field

    @device_guard
    def forward(self, seq):
        """Three (statistics, contraction) pairs: each LayerNorm + ReLU is applied as the following convolution loads
        its operand, the last contraction's store adds `seq`."""
        from .. import bytenet_tc
        if bytenet_tc.eligible(self, seq):        # bf16 inference: the tcgen05 form (NLC, dense2 contractions)
            return bytenet_tc.relu_block_forward(self, seq)
        st = self.stack
        seq = seq.contiguous()
        h = WF.fused_conv(seq, st[2].weight, st[2].bias, [0], ln=st[0])
        h = WF.fused_conv(h, st[5].conv1d.weight, st[5].conv1d.bias, st[5].offsets, ln=st[3])
        return WF.fused_conv(h, st[8].weight, st[8].bias, [0], ln=st[6], residual=seq)
