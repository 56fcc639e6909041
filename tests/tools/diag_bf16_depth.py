"""bf16 error vs depth: tensor-core path vs torch's own bf16 evaluation (CPU), both against the fp32 oracle on the
same bf16-rounded weights/input (diagnostic for the bf16 tolerance; SURVEY 7 hard part 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import wavenet_speech_b200 as W
from oracle import wavenet_oracle as O
from tests import _golden as G

C = 256
for depth in (1, 2, 4, 8, 15):
    torch.manual_seed(depth)
    dil = ([1, 2, 4, 8, 16] * 3)[:depth]
    layers = [(C, C, 2, d) for d in dil]
    net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False)
    sd = {k: v.detach().bfloat16().float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 1, 600).bfloat16().float()
    ref = O.raw_ctcnet_forward(sd, x, layers, softmax=False)
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    t16 = O.raw_ctcnet_forward(sd16, x.bfloat16(), layers, softmax=False).float()
    with torch.no_grad():
        y = net.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu()
    am = lambda a: (a.argmax(1) == ref.argmax(1)).float().mean().item()
    print("blocks %2d (+input block): ours linf %.3e l2 %.3e argmax %.4f | torch-bf16 linf %.3e l2 %.3e argmax %.4f" % (
        depth, G.rel_linf(y, ref), G.rel_l2(y, ref), am(y), G.rel_linf(t16, ref), G.rel_l2(t16, ref), am(t16)))
