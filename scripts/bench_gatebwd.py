"""A/B of the d(gate) contraction with and without the gate-backward epilogue at the config-3 layer shape
(B = 32, T = 16383, C = 256): CUDA-event time of each variant.  `python scripts/bench_gatebwd.py [reps]`."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from wavenet_speech_b200 import fastpath as FP, training as TR

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, T, C = 32, 16383, 256
torch.manual_seed(0)
dres = torch.randn(B, T, C, device="cuda").bfloat16()
dsk = torch.randn(B, T, C, device="cuda").bfloat16()
w = (torch.randn(C, 2 * C, device="cuda") / 22).bfloat16()
sg = torch.sigmoid(torch.randn(B, T, C, device="cuda")).bfloat16()
gate = (torch.tanh(torch.randn(B, T, C, device="cuda")) * sg.float()).bfloat16()
zb = torch.zeros(C, device="cuda")
cs = torch.zeros(2 * C, device="cuda")


def fused():
    return FP.dense(dres, [0], w, zb, C, x2=dsk, offsets2=[0], gate_bwd=(gate, sg), colsum=cs)


def split():
    dg = FP.dense(dres, [0], w, zb, C, x2=dsk, offsets2=[0])
    return TR.gate_bwd_nlc(dg, gate, sg, want_bias=True, th_is_gate=True)[0]


def chunked(n):
    def run():
        dab = torch.empty((B, T, 2 * C), dtype=torch.bfloat16, device="cuda")
        for i in range(0, B, n):
            dg = FP.dense(dres[i:i + n], [0], w, zb, C, x2=dsk[i:i + n], offsets2=[0])
            TR.gate_bwd_nlc(dg, gate[i:i + n], sg[i:i + n], want_bias=True, th_is_gate=True, out=dab[i:i + n], dbias=cs)
        return dab
    return run


def plain():
    return FP.dense(dres, [0], w, zb, C, x2=dsk, offsets2=[0])


out = {}
for name, fn in (("fused", fused), ("two_launches", split), ("contraction_only", plain),
                 ("two_launches_chunks_of_8", chunked(8)), ("two_launches_chunks_of_4", chunked(4))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / reps
print(json.dumps(out))
