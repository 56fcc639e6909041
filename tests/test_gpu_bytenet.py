"""ByteNet residual blocks (reference modules/block.py:86-173) and the frame-at-a-time LinearConv1d
(modules/linear_conv_ops.py:39-68) on the CUDA path, through the C-ABI, against fixtures written by the REFERENCE's own
modules (oracle/gen_golden_bytenet.py: outputs and autograd gradients) and against the CPU oracle on seeded inputs.
fp32: <= 1e-5 relative on outputs, <= 2e-4 on gradients; bf16 storage: <= 2e-2."""
import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import _lib, ops
from wavenet_speech_b200 import functional as WF
from oracle import wavenet_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
GRAD_TOL = 2e-4


def _block(name, meta):
    cls = W.ResidualMUBlock if "_mu_" in name else W.ResidualReLUBlock
    return cls(meta["nchannels"], meta["k"], meta["d"])


def _err(a, ref):
    return G.rel_linf(a.detach().float().cpu(), ref)


def _grad_err(a, ref, scale):
    """|a - ref| over max(|ref|, 1e-3 * the largest gradient of the module): a gradient that is a cancellation to round-off
    (the c4 fixtures normalise over TWO channels, where LayerNorm's output is +-0.707 whatever the input and nothing flows
    back through it) is compared on the scale of the gradients that are not."""
    a, ref = a.detach().double().cpu().flatten(), ref.double().flatten()
    return float((a - ref).abs().max() / max(float(ref.abs().max()), 1e-3 * scale, 1e-30))


@pytest.mark.parametrize("name", [n for n in G.names() if n.startswith("bytenet_")])
def test_blocks_against_reference_fixtures(name):
    """Inference (fused: LayerNorm + ReLU on the operand load, MU gate and residual add in the epilogues) and training
    (same outputs, gradients of sum(y * probe)) against what the reference module + torch autograd wrote."""
    g = G.load(name)
    net = _block(name, g["meta"]).cuda()
    net.load_state_dict(g["sd"], strict=True)
    assert net.receptive_field == g["meta"]["rf"]
    x = g["inp"]["x"].cuda()
    with torch.no_grad():
        y_inf = net(x)
    assert _err(y_inf, g["out"]["y"]) <= FP32_TOL, (name, _err(y_inf, g["out"]["y"]))
    xg = x.clone().requires_grad_(True)
    y = net(xg)
    assert _err(y, g["out"]["y"]) <= FP32_TOL
    (y * g["inp"]["probe"].cuda()).sum().backward()
    scale = max(float(v.abs().max()) for k, v in g["out"].items() if k.startswith("grad"))
    assert _grad_err(xg.grad, g["out"]["grad_x"], scale) <= GRAD_TOL, ("dx", _grad_err(xg.grad, g["out"]["grad_x"], scale))
    for n, p in net.named_parameters():
        ref = g["out"]["grad/" + n]
        assert p.grad is not None, n
        assert _grad_err(p.grad, ref, scale) <= GRAD_TOL, (n, _grad_err(p.grad, ref, scale))


@pytest.mark.parametrize("kind,nch,k,d,B,T", [("relu", 256, 3, 4, 3, 1003), ("mu", 256, 3, 4, 3, 1003),
                                               ("relu", 64, 2, 16, 2, 512), ("mu", 96, 2, 1, 2, 130),
                                               ("relu", 50, 5, 2, 1, 77), ("mu", 34, 4, 3, 2, 61)])
def test_blocks_against_the_oracle(kind, nch, k, d, B, T):
    """Wider blocks (half = 128 channels spans a full weight tile, odd channel counts, T not a multiple of 4, kernel
    width 5 = chained tap launches): inference and training forward against the oracle, gradients against autograd
    through the oracle."""
    torch.manual_seed(1000 + nch + T)
    cls = W.ResidualMUBlock if kind == "mu" else W.ResidualReLUBlock
    fn = O.residual_mu_block if kind == "mu" else O.residual_relu_block
    net = cls(nch, k, d)
    net.init()
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("gamma") or n.endswith("beta"):
                p.add_(torch.randn_like(p) * 0.2)
    sd = {kk: v.detach().clone().requires_grad_(True) for kk, v in net.state_dict().items()}
    x = torch.randn(B, nch, T)
    xr = x.clone().requires_grad_(True)
    ref = fn(sd, "", xr, d)
    probe = torch.randn_like(ref)
    (ref * probe).sum().backward()
    net = net.cuda()
    with torch.no_grad():
        y_inf = net(x.cuda())
    assert _err(y_inf, ref.detach()) <= FP32_TOL, _err(y_inf, ref.detach())
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    assert _err(y, ref.detach()) <= FP32_TOL
    (y * probe.cuda()).sum().backward()
    scale = max([float(v.grad.abs().max()) for v in sd.values()] + [float(xr.grad.abs().max())])
    assert _grad_err(xg.grad, xr.grad, scale) <= GRAD_TOL, ("dx", _grad_err(xg.grad, xr.grad, scale))
    for n, p in net.named_parameters():
        assert _grad_err(p.grad, sd[n].grad, scale) <= GRAD_TOL, (n, _grad_err(p.grad, sd[n].grad, scale))


def test_inference_launch_counts():
    """The fused forward: 6 launches per block (3 statistics + 3 contractions; MU block: 2 statistics + 4 contractions),
    none of them a torch op on the activations."""
    torch.manual_seed(3)
    x = torch.randn(2, 64, 256).cuda()
    for cls in (W.ResidualReLUBlock, W.ResidualMUBlock):
        net = cls(64, 2, 2).cuda()
        with torch.no_grad():
            net(x)
            n0 = _lib.launch_count
            net(x)
        assert _lib.launch_count - n0 == 6, (cls.__name__, _lib.launch_count - n0)


def test_bf16_storage():
    torch.manual_seed(5)
    for kind, fn in (("relu", O.residual_relu_block), ("mu", O.residual_mu_block)):
        cls = W.ResidualMUBlock if kind == "mu" else W.ResidualReLUBlock
        net = cls(128, 3, 2)
        net.init()
        sd = {k: v.detach().to(torch.bfloat16).float() for k, v in net.state_dict().items()}
        x = torch.randn(2, 128, 300).to(torch.bfloat16)
        ref = fn(sd, "", x.float(), 2)
        net = net.cuda().to(torch.bfloat16)
        with torch.no_grad():
            y = net(x.cuda())
        assert y.dtype == torch.bfloat16
        assert _err(y, ref) <= 2e-2, (kind, _err(y, ref))


def test_layernorm_relu_kernels():
    """Statistics, stand-alone LayerNorm + ReLU and its backward (dx, dgamma, dbeta) against autograd through the oracle's
    layernorm; T not a multiple of the 128-frame tile, C not a multiple of the 8 channel groups."""
    torch.manual_seed(7)
    for (B, C, T) in [(2, 6, 9), (3, 100, 257), (1, 256, 1024)]:
        x = torch.randn(B, C, T) * 2 + 0.5
        gamma = (torch.ones(1, C, 1) + 0.2 * torch.randn(1, C, 1))
        beta = 0.2 * torch.randn(1, C, 1)
        xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        ref = torch.relu(O.layernorm(xr, gr, br))
        probe = torch.randn_like(ref)
        (ref * probe).sum().backward()
        xs = x.cuda()
        stats = ops.ln_stats(xs, 1e-6)
        mean = x.mean(1)
        rinv = 1.0 / (x.std(1) + 1e-6)
        assert _err(stats[..., 0], mean) <= 1e-6 and _err(stats[..., 1], rinv) <= 1e-6
        g32, b32 = gamma.reshape(-1).cuda(), beta.reshape(-1).cuda()
        y = ops.ln_relu_fwd(xs, stats, g32, b32)
        assert _err(y, ref.detach()) <= FP32_TOL
        dx, dg, db = ops.ln_relu_bwd(xs, stats, g32, b32, 1e-6, probe.cuda())
        assert _err(dx, xr.grad) <= GRAD_TOL, _err(dx, xr.grad)
        assert _err(dg, gr.grad.reshape(-1)) <= GRAD_TOL and _err(db, br.grad.reshape(-1)) <= GRAD_TOL


def test_multiplicative_unit_fused_epilogue():
    """MultiplicativeUnit in one launch (no grad) == contraction + gate kernel (with grad) == oracle; channel counts
    around the 32-channel packing group, with and without the residual riding on the store."""
    torch.manual_seed(9)
    for (ndim, k, d, T) in [(6, 3, 2, 20), (32, 2, 1, 64), (33, 1, 1, 50), (80, 4, 3, 131)]:
        net = W.MultiplicativeUnit(ndim, k, dilation=d).cuda()
        sd = {kk: v.detach().cpu() for kk, v in net.state_dict().items()}
        x = torch.randn(2, ndim, T)
        ref = O.multiplicative_unit(sd, "", x, d)
        with torch.no_grad():
            y1 = net(x.cuda())
        y2 = net(x.cuda().requires_grad_(True))
        assert _err(y1, ref) <= FP32_TOL and _err(y2, ref) <= FP32_TOL, (ndim, _err(y1, ref), _err(y2, ref))
        res = torch.randn_like(ref)
        y3 = WF.multiplicative_unit_fused(x.cuda(), net.convs, net.gate1.offsets, residual=res.cuda())
        assert _err(y3, ref + res) <= FP32_TOL


def test_linear_frame_and_stream_fixtures():
    g = G.load("linearconv_k3_d2")
    m = g["meta"]
    conv = W.LinearConv1d(m["cin"], m["cout"], m["k"], dilation=m["d"]).cuda()
    conv.load_state_dict(g["sd"])
    with torch.no_grad():
        y = conv.linear(g["inp"]["frame"].cuda())
        y3 = conv.linear(g["inp"]["frame"].cuda(), keep_dims=True)
    assert _err(y, g["out"]["y"]) <= FP32_TOL and tuple(y3.shape) == tuple(y.shape) + (1,)
    yg = conv.linear(g["inp"]["frame"].cuda().requires_grad_(True))          # autograd route: same numbers
    assert _err(yg, g["out"]["y"]) <= FP32_TOL
    for name in [n for n in G.names() if n.startswith("linearconv_stream")]:
        g = G.load(name)
        m = g["meta"]
        conv = W.LinearConv1d(m["cin"], m["cout"], m["k"], dilation=m["d"]).cuda()
        conv.load_state_dict(g["sd"])
        assert conv.receptive_field == m["rf"]
        seq = g["inp"]["seq"].cuda()
        st = conv.stream(seq.shape[0])
        ys = torch.stack([st.push(seq[:, :, t].contiguous()) for t in range(seq.shape[2])], 2)
        assert _err(ys, g["out"]["y"]) <= FP32_TOL, (name, _err(ys, g["out"]["y"]))
        st.reset()                                                            # a second sequence through the same state
        ys2 = torch.stack([st.push(seq[:, :, t].contiguous()) for t in range(seq.shape[2])], 2)
        assert torch.equal(ys, ys2)


@pytest.mark.parametrize("cin,cout,k,d,N,T,dtype", [(512, 512, 3, 4, 16, 40, torch.float32),
                                                    (100, 37, 5, 3, 9, 50, torch.float32),
                                                    (64, 64, 1, 1, 3, 5, torch.float32),
                                                    (256, 128, 2, 7, 8, 30, torch.bfloat16)])
def test_stream_matches_the_causal_convolution(cin, cout, k, d, N, T, dtype):
    """push() over a sequence == the causal convolution of the whole sequence (conv_ops.py:39-44) == linear() on every
    window; batches above the kernel's 8-item register tile, widths around the ring's wrap (T > receptive field)."""
    torch.manual_seed(cin + k)
    conv = W.LinearConv1d(cin, cout, k, dilation=d)
    w, b = conv.weight.detach().clone(), conv.bias.detach().clone()
    seq = torch.randn(N, cin, T)
    tol = FP32_TOL
    if dtype == torch.bfloat16:
        w, seq, tol = w.to(dtype).float(), seq.to(dtype).float(), 2e-2
    ref = O.causal_conv1d(seq, w, b, d)
    conv = conv.cuda().to(dtype)
    sq = seq.cuda().to(dtype)
    st = conv.stream(N)
    ys = torch.stack([st.push(sq[:, :, t].contiguous()) for t in range(T)], 2)
    assert _err(ys, ref) <= tol, _err(ys, ref)
    rf = conv.receptive_field
    if T >= rf:
        with torch.no_grad():
            yl = conv.linear(sq[:, :, T - rf:].contiguous())
        assert _err(yl, ref[:, :, -1]) <= tol


@pytest.mark.parametrize("kind,nch,k,d,B,T", [("relu", 512, 3, 4, 2, 1003), ("mu", 512, 3, 4, 2, 1003),
                                               ("relu", 384, 2, 16, 3, 512), ("mu", 384, 2, 5, 1, 777),
                                               ("relu", 256, 1, 1, 2, 300), ("mu", 128, 3, 2, 4, 64),
                                               ("mu", 256, 2, 512, 1, 2048)])
def test_tensor_core_form(kind, nch, k, d, B, T):
    """bf16 inference on the tcgen05 form (bytenet_tc.py): against the fp32 oracle on the same bf16-rounded weights and
    input (<= 2e-2), and against the generic CUDA-core kernels on the same tensors; the launches are dense2 contractions."""
    from wavenet_speech_b200 import bytenet_tc
    torch.manual_seed(2000 + nch + T + k)
    cls = W.ResidualMUBlock if kind == "mu" else W.ResidualReLUBlock
    fn = O.residual_mu_block if kind == "mu" else O.residual_relu_block
    net = cls(nch, k, d)
    net.init()
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("gamma") or n.endswith("beta"):
                p.add_(torch.randn_like(p) * 0.2)
    sd = {kk: v.detach().to(torch.bfloat16).float() for kk, v in net.state_dict().items()}
    x = torch.randn(B, nch, T).to(torch.bfloat16)
    ref = fn(sd, "", x.float(), d)
    net = net.cuda().to(torch.bfloat16)
    xg = x.cuda()
    with torch.no_grad():
        assert bytenet_tc.eligible(net, xg)
    log = []
    orig = _lib.call

    def spy(name, *a):
        log.append(name)
        return orig(name, *a)
    _lib.call = spy
    try:
        with torch.no_grad():
            y = net(xg)
    finally:
        _lib.call = orig
    assert "wnb200_dense_fwd_tc" in log and "wnb200_taps_fwd_ex" not in log and "wnb200_taps_fwd" not in log, log
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(ref.shape)
    assert _err(y, ref) <= 2e-2, (kind, nch, _err(y, ref))
    bytenet_tc.ENABLED = False
    try:
        with torch.no_grad():
            y_gen = net(xg)
    finally:
        bytenet_tc.ENABLED = True
    assert _err(y, y_gen.float().cpu()) <= 2e-2, _err(y, y_gen.float().cpu())
    with torch.no_grad():
        assert not bytenet_tc.eligible(net, xg.float())          # fp32 keeps the fp32-accurate kernels
    assert not bytenet_tc.eligible(net, xg)                      # gradients requested: generic kernels (they have a backward)


def test_randomised_small_shapes_and_edges():
    """Random small configurations (odd channel counts around the tile sizes, T down to 1 and below the dilation, kernel
    widths 1..6 incl. chained tap launches): forward without gradients (fused epilogues) and with (contraction + gate
    kernels) against the oracle; empty batch."""
    import random
    rng = random.Random(1234)
    for it in range(14):
        kind = rng.choice(["relu", "mu"])
        nch = 2 * rng.choice([2, 3, 5, 8, 17, 31, 33, 48, 65])
        k = rng.choice([1, 2, 2, 3, 4, 6]) if kind == "relu" else rng.choice([1, 2, 3, 4])
        d = rng.choice([1, 2, 3, 7, 16])
        B = rng.choice([1, 2, 3])
        T = rng.choice([1, 2, 3, 5, 17, 64, 129, 200])
        torch.manual_seed(it)
        cls = W.ResidualMUBlock if kind == "mu" else W.ResidualReLUBlock
        fn = O.residual_mu_block if kind == "mu" else O.residual_relu_block
        net = cls(nch, k, d)
        net.init()
        sd = {kk: v.detach().clone() for kk, v in net.state_dict().items()}
        x = torch.randn(B, nch, T)
        ref = fn(sd, "", x, d)
        net = net.cuda()
        with torch.no_grad():
            y1 = net(x.cuda())
        y2 = net(x.cuda().requires_grad_(True))
        assert tuple(y1.shape) == tuple(ref.shape)
        assert _err(y1, ref) <= FP32_TOL and _err(y2, ref) <= FP32_TOL, (kind, nch, k, d, B, T, _err(y1, ref), _err(y2, ref))
    net = W.ResidualReLUBlock(8, 2, 2).cuda()
    with torch.no_grad():
        assert tuple(net(torch.zeros(0, 8, 16, device="cuda")).shape) == (0, 8, 16)
    net = W.ResidualMUBlock(8, 2, 2).cuda()
    with torch.no_grad():
        assert tuple(net(torch.zeros(0, 8, 16, device="cuda")).shape) == (0, 8, 16)
    conv = W.LinearConv1d(4, 6, 3, dilation=2).cuda()
    with torch.no_grad():
        assert tuple(conv.linear(torch.zeros(0, 4, conv.receptive_field, device="cuda")).shape) == (0, 6)
