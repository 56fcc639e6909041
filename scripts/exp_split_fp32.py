"""Experiment: an fp32-accurate contraction on the tcgen05 kernels by operand splitting.
x = x1 + x2 + x3, W = W1 + W2 + W3 (bf16 terms: 24 mantissa bits each side), y = sum_{i+j<=4} Wi xj: every product is
exact in the fp32 accumulator.  One wnb200_dense_fwd_tc launch with K = 9 C: three zero-offset "taps" over the channel-
concatenated [x1|x2|x3] tensor with weight segments [W1|W1|W1], [W2|W2|0], [W3|0|0].  Error against an fp64 product,
beside the fp32-FMA generic kernel and cuBLAS fp32 on the same operands.
    python scripts/exp_split_fp32.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP, functional as WF

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def split3(v):
    v1 = v.bfloat16()
    r = v - v1.float()
    v2 = r.bfloat16()
    v3 = (r - v2.float()).bfloat16()
    return v1, v2, v3


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for (B, C, T, scale) in [(2, 256, 4096, 1.0), (2, 256, 4096, 300.0), (8, 256, 16384, 1.0)]:
    torch.manual_seed(0)
    x = (torch.randn(B, C, T, device="cuda") * scale)
    w = torch.empty(C, C, device="cuda")
    torch.nn.init.kaiming_uniform_(w)
    bias = torch.zeros(C, device="cuda")
    ref = torch.einsum("mc,bct->bmt", w.double(), x.double())
    den = float(ref.abs().max())
    out = {"B": B, "C": C, "T": T, "scale": scale}
    # generic fp32-FMA kernel and cuBLAS fp32
    y_ffma = WF.conv_taps(x, w.unsqueeze(2), bias, [0])
    out["err_ffma_kernel"] = float((y_ffma.double() - ref).abs().max()) / den
    y_blas = torch.einsum("mc,bct->bmt", w, x)
    out["err_cublas_fp32"] = float((y_blas.double() - ref).abs().max()) / den
    out["ms_ffma_kernel"] = round(timed(lambda: WF.conv_taps(x, w.unsqueeze(2), bias, [0])), 3)
    # split operands on the tensor-core kernel
    xs = split3(x)
    x_nlc = torch.cat([t.permute(0, 2, 1) for t in xs], 2).contiguous()          # [B, T, 3C] bf16
    w1, w2, w3 = split3(w)
    z = torch.zeros_like(w1)
    for name, segs in (("6_products", [[w1, w1, w1], [w2, w2, z], [w3, z, z]]),
                       ("3_products", [[w1, w1, z], [w2, z, z]]),
                       ("1_product", [[w1, z, z]])):
        wbig = torch.cat([torch.cat(s, 1) for s in segs], 1).contiguous()        # [C, ntaps * 3C]
        y = torch.empty(B, C, T, device="cuda")
        f = lambda: FP.dense(x_nlc, [0] * len(segs), wbig, bias, C, mode=1, out=y, n_out=C)
        f()
        out["err_tc_" + name] = float((y.double() - ref).abs().max()) / den
        out["ms_tc_" + name] = round(timed(f), 3)
    print(json.dumps(out), flush=True)
