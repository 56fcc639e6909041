"""Layer-by-layer fp32 error growth of the CUDA path vs the fp32 oracle, both measured against an fp64 oracle
(diagnostic for the deep-stack tolerance; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import wavenet_speech_b200 as W
from wavenet_speech_b200 import functional as WF
from oracle import wavenet_oracle as O
from tests import _golden as G

g = G.load("wavenet_test_shape")
m = g["meta"]
sd = g["sd"]
sd64 = {k: v.double() for k, v in sd.items()}
x = g["inp"]["x"]
net = W.WaveNet(m["in_dim"], m["entry_kwidth"], m["layers"], m["out_dim"], softmax=True).cuda()
net.load_state_dict(sd)


def stack(sdx, xx, n):
    out = O.causal_conv1d(xx, sdx["entry_conv1d.conv1d.weight"], sdx["entry_conv1d.conv1d.bias"], 1)
    outs = []
    for l, (_a, _b, _k, d) in enumerate(m["layers"][:n]):
        out, skip = O.residual_block(sdx, "convolutions.%d." % l, out, d, True)
        outs.append(out)
    return outs


o64 = stack(sd64, x.double(), 40)
o32 = stack(sd, x, 40)
with torch.no_grad():
    out = net.entry_conv1d(x.cuda())
    for l, blk in enumerate(net.convolutions):
        out, skip = blk(out)
        if l % 4 == 3 or l < 3:
            print("layer %2d  cuda-vs-fp64 %.2e   cpu32-vs-fp64 %.2e   |out|max %.2f" % (
                l, G.rel_linf(out.cpu().double(), o64[l]), G.rel_linf(o32[l].double(), o64[l]), o64[l].abs().max()))
