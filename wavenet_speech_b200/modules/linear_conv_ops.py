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
        out = WF.conv_taps(frame, self.weight, self.bias, self._ker_ixs, T_out=1)
        return out if keep_dims else out.squeeze(2)

    @device_guard
    def forward(self, in_seq):
        k, d, p = self.kernel_size[0], self.dilation[0], self.padding[0]
        t_out = in_seq.size(2) + 2 * p - d * (k - 1)
        return WF.conv_taps(in_seq, self.weight, self.bias, [j * d - p for j in range(k)], T_out=t_out)
