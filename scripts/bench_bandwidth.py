"""Achieved HBM bandwidth of the bandwidth-bound kernels (B1 featuriser, B2 mean-pool, B3 softmax / x-ent,
B4 LayerNorm, layout changes, gate backward, column sums) against the measured copy peak in MEASURED_PEAKS.json.
One JSON line per kernel: algorithmic bytes (each input read once, each output written once) / CUDA-event time.
Shapes are those of BASELINE config 2 (32 x 256 x 16384) unless noted; every tensor exceeds the 126 MB L2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP, ops, training as TR, _lib

peak = 6555.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, nbytes, fn, note=""):
    ms = timed(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GB/s": round(gbs, 1), "frac_of_copy_peak": round(gbs / peak, 3), "peak_GB/s": peak, "note": note}),
          flush=True)


B, C, T = 32, 256, 16384
n = B * C * T
bf, f32 = torch.bfloat16, torch.float32
x_ncl = torch.randn(B, C, T, device="cuda", dtype=bf)
x_nlc = torch.randn(B, T, C, device="cuda", dtype=bf)
x_f32 = torch.randn(B, T, C, device="cuda", dtype=f32)
d2, d3 = torch.randn_like(x_nlc), torch.randn_like(x_nlc)
tgt = torch.randint(0, C, (B, T), device="cuda")

# what a READ-ONLY pass reaches on this GPU with stock kernels (the copy peak counts a read and a write stream): torch's own
# linear sum and its channel reduction over the same NCL tensor
report("reference: torch.sum over 268 MB bf16 (read-only, linear)", 2 * n, lambda: x_ncl.sum())
report("reference: torch.amax over the channels of the NCL tensor (read-only)", 2 * n, lambda: x_ncl.amax(1))
report("ncl_to_nlc_bf16 (input layout change)", 4 * n, lambda: FP.ncl_to_nlc_bf16(x_ncl))
report("nlc_to_ncl (output layout change)", 4 * n, lambda: FP.nlc_to_ncl(x_nlc, bf))
report("leaky_to_bf16 (fp32 skip sum -> head input)", 6 * n, lambda: FP.leaky_to_bf16(x_f32))
report("softmax_fwd NCL bf16 (B3)", 4 * n, lambda: ops.softmax_fwd(x_ncl))
report("xent_fwd NCL bf16: log-softmax + NLL (B3)", 2 * n + 16 * B * T, lambda: ops.xent_fwd(x_ncl, tgt))
lse = ops.xent_fwd(x_ncl, tgt)[1]
gs = torch.ones(1, device="cuda")
report("xent_bwd NCL bf16", 4 * n + 12 * B * T, lambda: ops.xent_bwd(x_ncl, tgt, lse, gs))
report("argmax_channels NCL bf16", 2 * n + 8 * B * T, lambda: ops.argmax_channels(x_ncl))
# odd T (the train step works on T-1 = 16383 frames): rows are not 16-byte aligned -> element-wise tile path
xo, tgo = x_ncl[:, :, :T - 1].contiguous(), tgt[:, :T - 1].contiguous()
no = xo.numel()
report("xent_fwd NCL bf16, T = 16383 (unaligned rows)", 2 * no + 16 * B * (T - 1), lambda: ops.xent_fwd(xo, tgo))
lseo = ops.xent_fwd(xo, tgo)[1]
report("xent_bwd NCL bf16, T = 16383 (unaligned rows)", 4 * no + 12 * B * (T - 1), lambda: ops.xent_bwd(xo, tgo, lseo, gs))
report("argmax_channels NCL bf16, T = 16383 (unaligned rows)", 2 * no + 8 * B * (T - 1), lambda: ops.argmax_channels(xo))
h0 = torch.empty((B, T // 3, C), dtype=bf, device="cuda")
report("avgpool(3) NCL -> NLC bf16 (B2)", 2 * n + 2 * (n // 3),
       lambda: _lib.call("wnb200_avgpool_ncl_to_nlc", 1, B, C, T, 3, ops._p(x_ncl), 0, ops._p(h0), ops._stream()))
report("avgpool_fwd(3) NCL bf16 (B2, generic path)", 2 * n + 2 * (n // 3), lambda: ops.avgpool_fwd(x_ncl, 3))
g = torch.ones(C, device="cuda")
report("layernorm_fwd NCL bf16 (B4)", 4 * n + 8 * B * T, lambda: ops.layernorm_fwd(x_ncl, g, g, 1e-6))
report("gate_bwd_nlc (3 reads, 2C write)", 10 * n, lambda: TR.gate_bwd_nlc(x_nlc, d2, d3))
report("colsum_nlc (bias gradients)", 2 * n, lambda: TR.colsum(x_nlc))
report("leaky_bwd bf16", 6 * n, lambda: ops.leaky_bwd(x_nlc, d2))
# B1: raw-signal featuriser, ecoli RawCTCNet shape (fk = 3, F = 256): reads 4 B, writes F*2 B per sample
Bf, Tf, F, fk = 256, 4000, 256, 3
raw = torch.randn(Bf, 1, Tf, device="cuda", dtype=bf)
w0, b0 = torch.randn(F, fk, device="cuda"), torch.randn(F, device="cuda")
hf = torch.empty((Bf, Tf + fk - 1, F), dtype=bf, device="cuda")
report("featurize_nlc (B1: Conv1d(1,F,3) + LeakyReLU -> NLC bf16)", Bf * Tf * 2 + hf.numel() * 2,
       lambda: _lib.call("wnb200_featurize_nlc", 1, Bf, Tf, F, fk, ops._p(raw), ops._p(w0), ops._p(b0), 0, ops._p(hf),
                         ops._stream()), note="batch 256 x 4000")
logits = torch.randn(1024, 5, 4002, device="cuda", dtype=bf)
report("ctc_greedy_decode (argmax + collapse + pack)", logits.numel() * 2 + 1024 * 4002 * 2,
       lambda: ops.ctc_greedy_decode(logits), note="1024 reads x 4002 frames x 5 classes; one CTA per read")

# ---- ByteNet pieces: LayerNorm + ReLU on NLC rows (tensor-core form), the MU gate on rows, the way back to
# NCL with the residual add; the NCL statistics / apply / backward kernels of the generic form
import ctypes
from wavenet_speech_b200 import functional as WF
g32 = torch.ones(C, device="cuda") + 0.1 * torch.randn(C, device="cuda")
b32 = 0.1 * torch.randn(C, device="cuda")
y_nlc = torch.empty_like(x_nlc)
report("lnrelu_rows NLC bf16: LayerNorm + ReLU over contiguous channels (B4, tensor-core form)", 4 * n,
       lambda: _lib.call("wnb200_lnrelu_rows", B * T, C, ops._p(x_nlc), ops._p(g32), ops._p(b32), 1e-6, ops._p(y_nlc),
                         ops._stream()))
pre4 = torch.randn(B, T, 4 * C, device="cuda", dtype=bf)
report("mu_gate_rows NLC bf16: g1 tanh(g2 h + g3 tanh(u))", 2 * n * 6,
       lambda: _lib.call("wnb200_mu_gate_rows", B * T, C, *[ctypes.c_void_p(pre4.data_ptr() + 2 * C * u) for u in range(4)],
                         4 * C, ops._p(x_nlc), ops._p(y_nlc), ops._stream()))
out_ncl = torch.empty_like(x_ncl)
parts = (ctypes.c_void_p * 1)(x_nlc.data_ptr())
report("nlc_parts_to_ncl_add: NLC -> NCL + residual", 6 * n,
       lambda: _lib.call("wnb200_nlc_parts_to_ncl_add", B, C, T, 1, C, parts, ops._p(x_ncl), ops._p(out_ncl), ops._stream()))
report("ln_stats NCL bf16: per-frame mean / 1/(std+eps)", 2 * n + 8 * B * T, lambda: ops.ln_stats(x_ncl, 1e-6))
st = ops.ln_stats(x_ncl, 1e-6)
report("ln_relu_fwd NCL bf16 (given the statistics)", 4 * n + 8 * B * T, lambda: ops.ln_relu_fwd(x_ncl, st, g32, b32))
dy = torch.randn_like(x_ncl)
report("ln_relu_bwd NCL bf16: dx, dgamma, dbeta", 6 * n + 8 * B * T, lambda: ops.ln_relu_bwd(x_ncl, st, g32, b32, 1e-6, dy))
