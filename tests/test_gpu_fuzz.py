"""Randomised shapes for the fused residual-block kernel (CTA pair) against the oracle: every combination of channel
count, kernel width, dilation, causality, read length and batch drawn below goes through the three ways the kernel is
used -- plain (res + running skip sum), last layer of an inference stack (emits LeakyReLU(skip sum) as bf16) and the
training forward (keeps the gate and its two factors).  Seeds are fixed: the cases are the same on every run."""
import random

import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP
from oracle import wavenet_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def r16(t):
    return t.detach().bfloat16().float()


def rel(y, ref):
    return G.rel_linf(y.detach().float().cpu(), ref)


def _cases(n):
    rng = random.Random(20261018)
    out = []
    for i in range(n):
        C = rng.choice([128, 256])
        k = rng.choice([1, 2, 2, 3])
        d = rng.choice([1, 2, 3, 5, 8, 17, 64, 300, 512])
        causal = rng.random() < 0.5
        T = rng.choice([1, 7, 127, 128, 129, 255, 256, 257, 300, 640, 1000, 1499])
        B = rng.choice([1, 2, 3])
        out.append((i, C, k, d, causal, T, B))
    return out


@pytest.mark.parametrize("i,C,k,d,causal,T,B", _cases(20))
def test_block_kernel_random_shapes(i, C, k, d, causal, T, B):
    torch.manual_seed(1000 + i)
    blk = W.ResidualBlock(C, C, k, d, causal=causal)
    bn = torch.nn.Conv1d(C, C, 1)
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    sd = {kk: r16(v) if v.dim() > 1 else v.detach() for kk, v in blk.state_dict().items()}
    x = r16(torch.randn(B, C, T))
    res_ref, skip_ref = O.residual_block(sd, "", x, d, causal)
    contrib_ref = torch.nn.functional.conv1d(skip_ref, bn.weight.detach(), bn.bias.detach())
    pk = {kk: (v.cuda() if torch.is_tensor(v) else v) for kk, v in FP.pack_block(blk, bn).items()}
    xn = FP.ncl_to_nlc_bf16(x.cuda())
    prev = torch.randn(B, T, C).cuda()

    # plain: res + accumulate into the running skip sum
    res, skips = torch.empty_like(xn), prev.clone()
    FP.resblock(xn, pk, res, skips, False)
    torch.cuda.synchronize()
    assert rel(res.float().permute(0, 2, 1), res_ref) <= BF16_TOL
    assert rel((skips - prev).permute(0, 2, 1), contrib_ref) <= BF16_TOL

    # last layer of an inference stack: bf16 LeakyReLU(running sum + contribution), bit-identical to the two-pass way
    skips2 = prev.clone()
    FP.resblock(xn, pk, None, skips2, False)
    want = FP.leaky_to_bf16(skips2)
    got = torch.empty_like(xn)
    FP.resblock(xn, pk, None, prev.clone(), False, skips_act=got)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    first = torch.empty_like(xn)                        # ... and as the only layer (nothing to load)
    FP.resblock(xn, pk, None, torch.empty_like(prev), True, skips_act=first)
    only = torch.empty_like(prev)
    FP.resblock(xn, pk, None, only, True)
    assert torch.equal(first, FP.leaky_to_bf16(only))

    # training forward: same res / skips, plus the gate and its factors
    act, th, sg = torch.empty_like(xn), torch.empty_like(xn), torch.empty_like(xn)
    res3, skips3 = torch.empty_like(xn), prev.clone()
    FP.resblock(xn, pk, res3, skips3, False, save=(act, th, sg))
    torch.cuda.synchronize()
    assert torch.equal(res3, res) and torch.equal(skips3, skips)
    conv = O.causal_conv1d if causal else O.noncausal_conv1d
    a = conv(x, sd["conv_tanh.conv1d.weight"], sd["conv_tanh.conv1d.bias"], d)
    b = conv(x, sd["conv_sigmoid.conv1d.weight"], sd["conv_sigmoid.conv1d.bias"], d)
    assert (th.float().cpu().permute(0, 2, 1) - torch.tanh(a)).abs().max() <= 1.5e-2
    assert (sg.float().cpu().permute(0, 2, 1) - torch.sigmoid(b)).abs().max() <= 1.5e-2
    assert (act.float().cpu().permute(0, 2, 1) - torch.tanh(a) * torch.sigmoid(b)).abs().max() <= 1.5e-2


def _dense_cases(n):
    rng = random.Random(77)
    out = []
    for i in range(n):
        Cin = rng.choice([64, 128, 256])
        N = rng.choice([64, 128, 256])
        k = rng.choice([1, 2, 3])
        offs = sorted(rng.sample([-600, -64, -9, -2, -1, 0, 1, 3, 8, 130], k))
        two = rng.random() < 0.5
        Cin2 = rng.choice([64, 128, 256])
        offs2 = sorted(rng.sample([-5, -1, 0, 2, 40], rng.choice([1, 2])))
        T = rng.choice([1, 31, 128, 129, 256, 300, 777, 1024])
        B = rng.choice([1, 2, 3])
        out.append((i, Cin, N, tuple(offs), two, Cin2, tuple(offs2), T, B, rng.random() < 0.5))
    return out


def _shifted(x, o):
    """x[b, c, t + o] with zeros outside [0, T)."""
    T = x.shape[2]
    y = torch.zeros_like(x)
    lo, hi = max(0, -o), min(T, T - o)
    if hi > lo:
        y[:, :, lo:hi] = x[:, :, lo + o:hi + o]
    return y


@pytest.mark.parametrize("i,Cin,N,offs,two,Cin2,offs2,T,B,leaky", _dense_cases(16))
def test_dense_kernel_random_shapes(i, Cin, N, offs, two, Cin2, offs2, T, B, leaky):
    """CTA-pair dense contraction: one or two NLC sources, up to three taps each at arbitrary frame offsets (zero fill
    outside the read), optional LeakyReLU, optional fused column sums of the output."""
    torch.manual_seed(500 + i)
    x = r16(torch.randn(B, Cin, T))
    x2 = r16(torch.randn(B, Cin2, T)) if two else None
    K = Cin * len(offs) + (Cin2 * len(offs2) if two else 0)
    w = r16(torch.randn(N, K) / K ** 0.5)
    bias = torch.randn(N) * 0.1
    ref = bias.view(1, N, 1).expand(B, N, T).clone()
    col = 0
    for o in offs:
        ref += torch.einsum("nc,bct->bnt", w[:, col:col + Cin], _shifted(x, o))
        col += Cin
    if two:
        for o in offs2:
            ref += torch.einsum("nc,bct->bnt", w[:, col:col + Cin2], _shifted(x2, o))
            col += Cin2
    if leaky:
        ref = torch.nn.functional.leaky_relu(ref, 0.01)
    cs = torch.zeros(N, device="cuda")
    y = FP.dense(FP.ncl_to_nlc_bf16(x.cuda()), list(offs), w.cuda().bfloat16(), bias.cuda(), N, leaky=int(leaky),
                 x2=FP.ncl_to_nlc_bf16(x2.cuda()) if two else None, offsets2=list(offs2) if two else (), colsum=cs)
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 2, 1), ref) <= 1e-2
    want_cs = y.float().sum((0, 1))                       # sums of the bf16-rounded outputs, as a colsum launch would see
    assert torch.allclose(cs, want_cs, rtol=1e-4, atol=1e-3 * max(1.0, float(want_cs.abs().max())))


def _net_cases(n):
    rng = random.Random(4242)
    out = []
    for i in range(n):
        kind = rng.choice(["wavenet", "rawctc", "classifier"])
        C = rng.choice([128, 256])
        nl = rng.choice([1, 2, 3, 4])
        layers = [(C, C, rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 4, 7, 16, 33])) for _ in range(nl)]
        T = rng.choice([5, 64, 130, 257, 500, 999])
        B = rng.choice([1, 2, 3])
        out.append((i, kind, C, tuple(layers), T, B, rng.random() < 0.5, rng.choice([1, 2, 3, 4]), rng.random() < 0.5))
    return out


@pytest.mark.parametrize("i,kind,C,layers,T,B,softmax,aux,causal", _net_cases(12))
def test_networks_random_configs(i, kind, C, layers, T, B, softmax, aux, causal):
    """The three drop-in networks on the tensor-core path, random layer lists / lengths / batch sizes, against the oracle
    evaluated on the same bf16-rounded weights and inputs (logits within 2e-2 of the largest |logit|)."""
    torch.manual_seed(9000 + i)
    layers = [tuple(l) for l in layers]
    if kind == "wavenet":
        net = W.WaveNet(C, 2, layers, C, softmax=softmax)
        lev = torch.randint(0, C, (B, T))
        x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0)
        ref_fn = lambda sd: O.wavenet_forward(sd, x, layers, softmax=softmax)
    elif kind == "rawctc":
        net = W.RawCTCNet(C, aux, 5, layers, C, softmax=softmax, causal=causal)      # aux = feature kernel width
        x = r16(torch.randn(B, 1, T))
        ref_fn = lambda sd: O.raw_ctcnet_forward(sd, x, layers, softmax=softmax, causal=causal)
    else:
        T = max(T, aux)                                                               # aux = pool width
        net = W.WaveNetClassifier(C, 5, layers, C, pool_kernel_size=aux, softmax=softmax)
        x = r16(torch.randn(B, C, T))
        ref_fn = lambda sd: O.classifier_forward(sd, x, layers, pool_kernel_size=aux, softmax=softmax)
    ref = ref_fn({k: r16(v) for k, v in net.state_dict().items()})
    net = net.cuda().bfloat16().eval()
    with torch.no_grad():
        net(x.cuda().bfloat16())                     # first call packs the weights (one pack launch per block, cached)
        before = W._lib.launch_count
        y = net(x.cuda().bfloat16())
    # the tensor-core pipeline ran (one fused launch per block), not the generic kernels (three per block)
    assert W._lib.launch_count - before <= len(layers) + 6
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel(y, ref) <= BF16_TOL
