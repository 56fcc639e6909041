"""Precise format: output error / greedy-decode agreement and forward time versus the number of LAST blocks of a stack
whose stream carries its fp16 hi half only (fastpath.HI_ONLY_TAIL).  One JSON line per setting.  Run on a GPU box:
    python tests/tools/parity_hi_only_tail.py > profiles/r2_parity_hi_only_tail.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP
from wavenet_speech_b200.utils import signal_gen as SG
from oracle import wavenet_oracle as O
from tests import _golden as G

C = 256
r16 = lambda t: t.detach().bfloat16().float()
DIL = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2


def metrics(y, ref):
    y = y.float().cpu()
    return {"rel_linf": G.rel_linf(y, ref), "argmax_agree": float((y.argmax(1) == ref.argmax(1)).float().mean())}


def case(kind, seed):
    torch.manual_seed(seed)
    if kind == "wavenet20":
        layers = [(C, C, 2, d) for d in DIL]
        net = W.WaveNet(C, 2, layers, C, softmax=False)
        lev = torch.from_numpy(SG.quantized_batch(2, 1500, seed=seed + 50))
        x = torch.zeros(2, C, 1500).scatter_(1, lev.unsqueeze(1), 1.0)
        fwd = lambda sd_, x_: O.wavenet_forward(sd_, x_, layers, softmax=False)
    else:
        layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3]
        net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False)
        x = r16(torch.from_numpy(SG.raw_batch(2, 2000, seed=seed + 60)))
        fwd = lambda sd_, x_: O.raw_ctcnet_forward(sd_, x_, layers, softmax=False)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    return net.cuda().bfloat16().eval(), x.cuda().bfloat16(), fwd(sd, x)


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


cases = [(k, s, case(k, s)) for k in ("wavenet20", "rawctcnet16") for s in (1, 2)]
torch.manual_seed(0)
big = W.WaveNet(C, 2, [(C, C, 2, d) for d in DIL], C, softmax=True).cuda().bfloat16().eval()
lev = torch.from_numpy(SG.quantized_batch(8, 16384, seed=3)).repeat(4, 1)
xb = torch.zeros(32, C, 16384, dtype=torch.bfloat16, device="cuda").scatter_(1, lev.cuda().unsqueeze(1), 1.0)
tails = [int(a) for a in sys.argv[1:]] or [0, 6, 8, 10, 12, 14, 19]
with torch.no_grad():
    for rep in range(2):
        for tail in tails:
            FP.HI_ONLY_TAIL = tail
            out = {"hi_only_tail": tail, "rep": rep}
            for kind, seed, (net, x, ref) in cases:
                out["%s_seed%d" % (kind, seed)] = metrics(net(x), ref)
            out["config2_forward_ms"] = round(timed(lambda: big(xb)), 3)
            print(json.dumps(out), flush=True)
