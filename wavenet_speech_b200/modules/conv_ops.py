"""CausalConv1d / NonCausalConv1d drop-ins (reference: modules/conv_ops.py).

The reference pads an nn.Conv1d and slices the result back to the input length (conv_ops.py:31-34,43-44
and :65-68,78-79).  Here the same map is evaluated directly as a tap sum with zero fill at the true
sequence ends -- the padded columns are never computed.  `self.conv1d` is kept as an nn.Conv1d purely as
the parameter container, so the state_dict keys (`conv1d.weight`, `conv1d.bias`), shapes and default
initialisation are the reference's."""
import math

import torch
import torch.nn as nn

from .. import functional as WF
from ..ops import device_guard


def autopad(k, d):
    """Padding that keeps the temporal length for kernel k, dilation d (reference conv_ops.py:104-116)."""
    return WF.autopad(k, d)


def compute_new_length(seq_len, pad, dil, ker):
    """Length of a stride-1 Conv1d output (reference conv_ops.py:85-88)."""
    return float(math.floor(seq_len + 2 * pad - dil * (ker - 1)))


def reshape_in(seq):
    """(N, C, L) -> (N*L, C) plus the (N, L) needed to undo it (reference conv_ops.py:91-94).  Kept for
    callers; the kernels in this package never need it."""
    n, c, l = seq.size()
    return seq.permute(0, 2, 1).contiguous().view(n * l, c), (n, l)


def reshape_out(seq, dims):
    """Inverse of reshape_in (reference conv_ops.py:97-101)."""
    n, l = dims
    return seq.view(n, l, -1).permute(0, 2, 1).contiguous()


class _DilatedConv1d(nn.Module):
    causal = True

    def __init__(self, in_channels, out_channels, kernel_width, dilation=1):
        super(_DilatedConv1d, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_width = kernel_width
        self.dilation = dilation
        self.padding = (kernel_width - 1) * dilation if self.causal else autopad(kernel_width, dilation)
        # parameter container only (never called): same names / shapes / init as the reference
        self.conv1d = nn.Conv1d(in_channels, out_channels, kernel_width, stride=1, padding=self.padding,
                                dilation=dilation)
        self.receptive_field = kernel_width + (dilation - 1) * (kernel_width - 1)

    @property
    def offsets(self):
        return WF.tap_offsets(self.kernel_width, self.dilation, self.causal)

    @device_guard
    def forward(self, seq):
        return WF.conv_taps(seq, self.conv1d.weight, self.conv1d.bias, self.offsets)


class CausalConv1d(_DilatedConv1d):
    """y[t] = b + sum_j W[:,:,j] x[t - (k-1-j) d], zeros before the start of the sequence."""
    causal = True


class NonCausalConv1d(_DilatedConv1d):
    """y[t] = b + sum_j W[:,:,j] x[t + j d - autopad(k,d)], zero padded on both sides."""
    causal = False
