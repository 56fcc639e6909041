"""LinearConv1d (reference: modules/linear_conv_ops.py): an nn.Conv1d that can also evaluate a single
output frame from its receptive field (`linear`).  The reference's `forward` is an unimplemented stub
that returns None (linear_conv_ops.py:34-37); here it is the batched convolution the docstring promises,
i.e. the same tap sum as `linear` applied at every valid frame."""
import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard
from .conv_ops import autopad, compute_new_length, reshape_in, reshape_out  # re-exported like the reference


def get_ker_ixs(d, k):
    """Frame indices a width-k, dilation-d kernel reads inside its receptive field
    (reference linear_conv_ops.py:112-123)."""
    return list(range(0, k * d - (d - 1), d))


class ConvStream(object):
    """Frame-at-a-time state of one LinearConv1d: ring [receptive_field, N, C_in] on the device + a step counter."""

    def __init__(self, conv, batch_size, dtype, device):
        if conv.padding[0] != 0:
            raise NotImplementedError("ConvStream is the causal frame-at-a-time form; construct the conv with padding=0")
        if conv.kernel_size[0] > 32:
            raise NotImplementedError("ConvStream supports kernel widths up to 32")
        self.conv = conv
        self.hist = torch.zeros((conv.receptive_field, batch_size, conv.in_channels), dtype=dtype, device=device)
        self.step = 0

    def reset(self):
        self.hist.zero_()
        self.step = 0

    @torch.no_grad()
    def push(self, x_t):
        """x_t: (batch, in_channels) = input frame number `self.step` -> (batch, out_channels)."""
        c = self.conv
        w = c.weight if c.weight.dtype == x_t.dtype else c.weight.to(x_t.dtype)
        with torch.cuda.device(x_t.device):
            y = WF.ops.linear_step(x_t.contiguous(), w.contiguous(), WF._f32(c.bias), c.dilation[0], self.step, self.hist)
        self.step += 1
        return y


class LinearConv1d(nn.Conv1d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True):
        super(LinearConv1d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation,
                                           groups, bias)
        if stride != 1 or groups != 1:
            raise NotImplementedError("LinearConv1d kernels support stride=1, groups=1")
        self._ker_ixs = get_ker_ixs(dilation, kernel_size)

    @property
    def receptive_field(self):
        k, d = self.kernel_size[0], self.dilation[0]
        return k + (d - 1) * (k - 1)

    def linear(self, frame, keep_dims=False):
        """frame: (batch, in_channels, receptive_field) -> (batch, out_channels[, 1])."""
        assert frame.size(2) == self.receptive_field
        for hook in self._forward_pre_hooks.values():
            hook(self, frame)
        if WF.needs_grad(frame, self.weight, self.bias) or self.kernel_size[0] > 32:
            out = WF.conv_taps(frame, self.weight, self.bias, self._ker_ixs, T_out=1)
            return out if keep_dims else out.squeeze(2)
        # inference: the one-frame GEMV kernel (weights streamed once, gathered taps staged in shared memory)
        w = self.weight if self.weight.dtype == frame.dtype else self.weight.to(frame.dtype)
        out = WF.ops.linear_frame(frame, w.contiguous(), WF._f32(self.bias), self.dilation[0])
        return out.unsqueeze(2) if keep_dims else out

    def stream(self, batch_size, dtype=None, device=None):
        """Incremental evaluation (the use the reference's docstring names, linear_conv_ops.py:5-8): a `ConvStream` whose
        `push(x_t)` returns the output frame for input frame t, reading the k - 1 earlier taps from a device ring of the
        last `receptive_field` frames (zeros before the first frame = the causal padding, conv_ops.py:39-44).  A loop
        of `push` over a sequence equals `CausalConv1d` / `linear` on every window; per step it does k taps of work
        where re-evaluating `linear` on a sliding window re-reads the whole receptive field."""
        return ConvStream(self, batch_size, dtype or self.weight.dtype, device or self.weight.device)

    @device_guard
    def forward(self, in_seq):
        k, d, p = self.kernel_size[0], self.dilation[0], self.padding[0]
        t_out = in_seq.size(2) + 2 * p - d * (k - 1)
        return WF.conv_taps(in_seq, self.weight, self.bias, [j * d - p for j in range(k)], T_out=t_out)
