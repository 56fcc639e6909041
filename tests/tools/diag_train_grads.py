import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import wavenet_speech_b200 as W
from oracle import wavenet_oracle as O
from tests import _golden as G
from tests.test_gpu_tc_train import r16, _oracle_grads, _perturb_biases, _Stored, _stored_stack

def run(C, nl, T, B, softmax):
    torch.manual_seed(C + nl)
    layers = [(C, C, 2, 2 ** i) for i in range(nl)]
    net = W.WaveNet(C, 2, layers, C, softmax=softmax)
    _perturb_biases(net)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    lev = torch.randint(0, C, (B, T))
    x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0) + r16(torch.randn(B, C, T) * 0.05)
    x = r16(x)
    R = r16(torch.randn(B, C, T))
    net = net.cuda()
    xg = x.cuda().bfloat16().requires_grad_(True)
    y = net(xg)
    _x, _offs, saved, skips_act, h1, _out = y.grad_fn.keep
    stored = _Stored([saved[0][0]] + _stored_stack(saved, skips_act, h1))
    yref, gref, dxref = _oracle_grads(sd, lambda s, xx: O.wavenet_forward(s, xx, layers, softmax=softmax, q=stored), x, R)
    print("fwd", G.rel_linf(y.float().cpu(), yref))
    (y.float() * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    d = (xg.grad.float().cpu() - dxref).abs()
    i = int(d.argmax())
    print("dx linf", float(d.max() / dxref.abs().max()), "l2", G.rel_l2(xg.grad.float().cpu(), dxref), "at", 
          (i // (C * T), (i // T) % C, i % T), "got", float(xg.grad.float().cpu().flatten()[i]), "ref", float(dxref.flatten()[i]))
    # error by time position
    et = d.amax((0, 1))
    print("dx err by t (first 8, last 8):", [round(float(v), 3) for v in et[:8]], [round(float(v), 3) for v in et[-8:]])
    print("dx err top t:", torch.topk(et, 8).indices.tolist())
    for name, p in net.named_parameters():
        ref = gref[name]
        if ref is None:
            print("%-50s ref None, got %s" % (name, None if p.grad is None else float(p.grad.abs().max())))
            continue
        if p.grad is None:
            print("%-50s MISSING" % name); continue
        print("%-50s linf %.4f l2 %.4f  |ref|max %.3g" % (name, G.rel_linf(p.grad.float().cpu(), ref), G.rel_l2(p.grad.float().cpu(), ref), float(ref.abs().max())))

run(128, 3, 300, 2, False)
run(256, 2, 520, 2, False)
