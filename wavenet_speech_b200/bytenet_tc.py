"""Tensor-core (tcgen05) inference form of the ByteNet residual blocks (reference modules/block.py:86-173).

bf16 input, no gradients, C in {128, 256, 384, 512}, kernel width <= 3: the block runs in the NLC layout of the
tensor-core path -- every convolution is `wnb200_dense_fwd_tc` (CTA-pair tcgen05 contraction, taps staged by TMA, zero fill
= the causal padding), LayerNorm + ReLU is a row kernel over a frame's contiguous channels, the MultiplicativeUnit's gate
an elementwise row kernel, and the way back to NCL carries the block's `seq +`.  Products of bf16 operands are exact in
the fp32 accumulator, so this is the arithmetic of the generic kernels on bf16 storage up to the bf16 rounding of the
normalised operand (held to <= 2e-2 against an fp32 evaluation of the same bf16-rounded weights, tests/test_gpu_bytenet.py).
Anything else (fp32, gradients, other shapes) stays on the generic kernels (functional.fused_conv)."""
import ctypes

import torch

from . import _lib, fastpath, ops
from . import functional as WF

ENABLED = True


def eligible(block, seq):
    if not ENABLED or not torch.is_tensor(seq) or not seq.is_cuda or seq.dim() != 3 or seq.dtype != torch.bfloat16:
        return False
    if WF.needs_grad(seq, *block.parameters()):
        return False
    C = block.nchannels
    if C % 128 != 0 or C < 128 or C > 512 or seq.shape[1] != C or seq.shape[0] == 0 or seq.shape[2] == 0:
        return False
    if seq.shape[0] > 65535:
        return False
    return block.k_width <= 3


def _wmat(weight):
    """conv weight [N, Cin, k] (or [N, Cin]) -> bf16 [N, k * Cin], tap-major columns (wnb200_dense_t.w)."""
    def build():
        w = weight.detach()
        if w.dim() == 2:
            w = w.unsqueeze(2)
        return w.permute(0, 2, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
    return WF._cached_layout([weight], "tc_wmat", torch.bfloat16, build)


def _conv(x_nlc, conv, offsets):
    """One convolution (<= 256 output channels per launch; wider ones as column blocks) -> list of NLC tensors."""
    w, b = _wmat(conv.weight), WF._f32(conv.bias)
    N = w.shape[0]
    if N <= 256:
        return [fastpath.dense(x_nlc, offsets, w, b, N)]
    step = 256 if N % 256 == 0 else (128 if N % 128 == 0 else 64)
    return [fastpath.dense(x_nlc, offsets, w[n0:n0 + step], b[n0:n0 + step], step) for n0 in range(0, N, step)]


def _lnrelu(x_nlc, ln):
    if ln.dim != 1:
        raise NotImplementedError("LayerNorm kernel normalises dim=1 of a (B, C, T) tensor")
    B, T, C = x_nlc.shape
    y = torch.empty_like(x_nlc)
    _lib.call("wnb200_lnrelu_rows", B * T, C, ops._p(x_nlc), ops._p(WF._flat32(ln.gamma)), ops._p(WF._flat32(ln.beta)),
              float(ln.eps), ops._p(y), ops._stream())
    return y


def _mu(h_nlc, mu):
    """MultiplicativeUnit on NLC rows: the four convolutions (one launch when 4 * C <= 256, else one per unit) + the gate."""
    B, T, C = h_nlc.shape
    offs = mu.gate1.offsets
    convs = mu.convs
    if 4 * C <= 256:
        w = WF._cached_layout([c.weight for c in convs], "tc_mu_w", torch.bfloat16,
                              lambda: torch.cat([_wmat(c.weight) for c in convs], 0).contiguous())
        b = WF._cached_layout([c.bias for c in convs], "tc_mu_b", torch.float32,
                              lambda: torch.cat([c.bias.detach().float() for c in convs], 0).contiguous())
        pre = fastpath.dense(h_nlc, offs, w, b, 4 * C)
        ptrs = [pre.data_ptr() + 2 * C * u for u in range(4)]
        pitch, keep = 4 * C, [pre]
    else:
        keep = [fastpath.dense(h_nlc, offs, _wmat(c.weight), WF._f32(c.bias), C) for c in convs]
        ptrs = [t.data_ptr() for t in keep]
        pitch = C
    out = torch.empty_like(h_nlc)
    _lib.call("wnb200_mu_gate_rows", B * T, C, *[ctypes.c_void_p(p) for p in ptrs], pitch, ops._p(h_nlc), ops._p(out),
              ops._stream())
    return out


def _finish(parts, seq):
    """NLC column blocks -> NCL, plus `seq` (block.py:119,166)."""
    B, C, T = seq.shape
    out = torch.empty_like(seq)
    arr = (ctypes.c_void_p * len(parts))(*[p.data_ptr() for p in parts])
    _lib.call("wnb200_nlc_parts_to_ncl_add", B, C, T, len(parts), parts[0].shape[2], arr, ops._p(seq), ops._p(out),
              ops._stream())
    return out


def relu_block_forward(block, seq):
    ops.check_device()
    st = block.stack
    seq = seq.contiguous()
    x = fastpath.ncl_to_nlc_bf16(seq)
    h = _conv(_lnrelu(x, st[0]), st[2], [0])[0]
    h = _conv(_lnrelu(h, st[3]), st[5].conv1d, st[5].offsets)[0]
    parts = _conv(_lnrelu(h, st[6]), st[8], [0])
    return _finish(parts, seq)


def mu_block_forward(block, seq):
    ops.check_device()
    st = block.stack
    seq = seq.contiguous()
    x = fastpath.ncl_to_nlc_bf16(seq)
    h = _conv(_lnrelu(x, st[0]), st[2], [0])[0]
    h = _mu(_mu(_lnrelu(h, st[3]), st[5]), st[6])
    parts = _conv(h, st[7], [0])
    return _finish(parts, seq)
