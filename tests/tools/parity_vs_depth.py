"""Error and greedy-decode agreement versus DEPTH for the tensor-core path's two formats, torch's own bf16 evaluation
and the fp32 CUDA path -- all against the fp32 CPU oracle on the same bf16-rounded weights and input.
Writes one JSON document (default profiles/r2_parity_vs_depth.json).  Run on a GPU box:
    python tests/tools/parity_vs_depth.py [out.json]"""
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
DEPTHS = (1, 5, 10, 16, 20)
r16 = lambda t: t.detach().bfloat16().float()


def metrics(y, ref):
    y = y.float().cpu()
    return {"rel_linf": G.rel_linf(y, ref), "rel_l2": G.rel_l2(y, ref),
            "argmax_agree": float((y.argmax(1) == ref.argmax(1)).float().mean())}


def run(kind, depth, seed):
    torch.manual_seed(seed)
    if kind == "wavenet":
        dil = ([1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2)[:depth]
        layers = [(C, C, 2, d) for d in dil]
        net = W.WaveNet(C, 2, layers, C, softmax=False)
        B, T = 2, 1500
        lev = torch.from_numpy(SG.quantized_batch(B, T, seed=seed + 50))
        x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0)
        fwd = lambda sd_, x_: O.wavenet_forward(sd_, x_, layers, softmax=False)
    else:
        dil = ([1, 2, 4, 8, 16] * 4)[:depth - 1]          # + the input block = `depth` blocks
        layers = [(C, C, 2, d) for d in dil]
        net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False)
        x = r16(torch.from_numpy(SG.raw_batch(2, 2000, seed=seed + 60)))
        fwd = lambda sd_, x_: O.raw_ctcnet_forward(sd_, x_, layers, softmax=False)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    ref = fwd(sd, x)
    out = {"kind": kind, "blocks": depth, "seed": seed, "ref_absmax": float(ref.abs().max())}
    t16 = fwd({k: v.bfloat16() for k, v in sd.items()}, x.bfloat16())
    out["torch_bf16_cpu"] = metrics(t16, ref)
    with torch.no_grad():
        g32 = net.cuda()
        out["ours_fp32_generic"] = metrics(g32(x.cuda()), ref)
        g16 = g32.bfloat16().eval()
        out["ours_precise_f16x2"] = metrics(g16(x.cuda().bfloat16()), ref)
        with FP.tc_precision("fast"):
            out["ours_fast_bf16"] = metrics(g16(x.cuda().bfloat16()), ref)
    return out


if __name__ == "__main__":
    dest = sys.argv[1] if len(sys.argv) > 1 else os.path.join("profiles", "r2_parity_vs_depth.json")
    rows = []
    for kind in ("wavenet", "raw_ctcnet"):
        for depth in DEPTHS:
            if kind == "raw_ctcnet" and depth == 1:
                depth = 2                    # RawCTCNet always has the input block + at least one more
            for seed in (0, 1):
                r = run(kind, depth, seed)
                rows.append(r)
                print(json.dumps(r), flush=True)
    doc = {"what": "output error (rel L-inf over max|ref|, rel L2) and per-frame argmax agreement vs the fp32 CPU oracle on "
                   "the same bf16-rounded weights and input, by number of residual blocks; 256 channels, k = 2; "
                   "wavenet: causal, dilations (1..512) x 2 truncated, logits (softmax=False), 2 x 1500 one-hot frames; "
                   "raw_ctcnet: non-causal, input block + dilations (1,2,4,8,16) x n, 2 x 2000 samples",
           "tolerance_north_star": 2e-2, "rows": rows}
    with open(dest, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", dest)
