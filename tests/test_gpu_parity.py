"""Parity of the CUDA path (through the C-ABI, via the drop-in modules) against the reference's own outputs
(tests/golden/) and against the CPU oracle on seeded inputs.  Tolerances: fp32 <= 1e-5 relative (L-inf
relative to the largest |logit|, as BASELINE.json's north_star states) -- see DESIGN.md for why deep
stacks are additionally compared against an fp64 oracle."""
import pytest
import torch

import wavenet_speech_b200 as W
from oracle import wavenet_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5


def dev(t):
    return t.cuda()


def check(y, ref, tol=FP32_TOL, what=""):
    y = y.detach().float().cpu()
    assert y.shape == ref.shape, (what, y.shape, ref.shape)
    e = G.rel_linf(y, ref)
    assert e <= tol, "%s rel_linf %.3e > %.1e (rel_l2 %.3e)" % (what, e, tol, G.rel_l2(y, ref))


@pytest.mark.parametrize("name", [n for n in G.names() if "conv_k" in n and "linear" not in n])
def test_conv_ops(name):
    g = G.load(name)
    m = g["meta"]
    cls = W.CausalConv1d if m["causal"] else W.NonCausalConv1d
    net = cls(m["cin"], m["cout"], m["k"], dilation=m["d"]).cuda()
    net.load_state_dict(g["sd"])
    check(net(dev(g["inp"]["x"])), g["out"]["y"], what=name)


@pytest.mark.parametrize("name", ["block_causal", "block_noncausal", "block_noncausal_k3_d3"])
def test_residual_block(name):
    g = G.load(name)
    m = g["meta"]
    net = W.ResidualBlock(m["cin"], m["cout"], m["k"], m["d"], causal=m["causal"]).cuda()
    net.load_state_dict(g["sd"])
    res, skip = net(dev(g["inp"]["x"]))
    check(res, g["out"]["res"], what=name + ".res")
    check(skip, g["out"]["skip"], what=name + ".skip")
    assert tuple(res.shape) == tuple(g["out"]["res"].shape)     # reference tests/test_block.py:32-40


@pytest.mark.parametrize("name", ["wavenet_test_shape", "wavenet_onehot_c32"])
def test_wavenet(name):
    g = G.load(name)
    m = g["meta"]
    net = W.WaveNet(m["in_dim"], m["entry_kwidth"], m["layers"], m["out_dim"], softmax=m["softmax"]).cuda()
    net.load_state_dict(g["sd"])
    with torch.no_grad():
        y = net(dev(g["inp"]["x"]))
    # wavenet_test_shape is the reference's own test shape (tests/test_wavenet.py: 40 untrained layers on randn
    # input).  Its residual stream grows to ~1e7 and the REFERENCE's fp32 output is itself 7.9e-5 away from an
    # fp64 evaluation of the same weights (measured; a 1e-7 input perturbation moves it by 5e-5), so 1e-5 against
    # the fp32 golden is not defined for it: hold it to 1e-3 and to a small multiple of the reference's own
    # fp32 noise against the fp64 oracle.  The 6-layer one-hot fixture is held to the 1e-5 bar.
    deep = name == "wavenet_test_shape"
    check(y, g["out"]["y"], tol=1e-3 if deep else FP32_TOL, what=name)
    sd64 = {k: v.double() for k, v in g["sd"].items()}
    y64 = O.wavenet_forward(sd64, g["inp"]["x"].double(), m["layers"], softmax=m["softmax"])
    err_new = G.rel_linf(y.cpu().double(), y64)
    err_ref = G.rel_linf(g["out"]["y"].double(), y64)
    assert err_new <= max(8.0 * err_ref, 5e-6), (err_new, err_ref)


@pytest.mark.parametrize("name", ["rawctcnet_default", "rawctcnet_positions", "rawctcnet_causal",
                                  "rawctcnet_example_json"])
def test_raw_ctcnet(name):
    g = G.load(name)
    m = g["meta"]
    net = W.RawCTCNet(m["num_features"], m["feature_kwidth"], m["num_labels"], m["layers"], m["out_dim"],
                      positions=m["positions"], softmax=m["softmax"], causal=m["causal"]).cuda()
    net.load_state_dict(g["sd"])
    with torch.no_grad():
        y = net(dev(g["inp"]["x"]))
    check(y, g["out"]["y"], what=name)
    # greedy decode: per-frame argmax identical (sequence_decoders.py:21-23) and collapsed sequence identical
    am = W.ops.argmax_channels(y).cpu()
    ref_am = O.argmax_decode(g["out"]["y"].permute(0, 2, 1))
    if not m["softmax"] or True:
        agree = (am == ref_am).float().mean().item()
        assert agree == 1.0, agree
        for b in range(am.shape[0]):
            assert O.collapse_decode(am[b]) == O.collapse_decode(ref_am[b])


def test_classifier():
    g = G.load("classifier_pool3")
    m = g["meta"]
    net = W.WaveNetClassifier(m["in_dim"], m["num_labels"], m["layers"], m["out_dim"],
                              pool_kernel_size=m["pool_kernel_size"], softmax=m["softmax"]).cuda()
    net.load_state_dict(g["sd"])
    with torch.no_grad():
        y = net(dev(g["inp"]["x"]))
    check(y, g["out"]["y"], what="classifier")


def test_layernorm_linearconv_mu():
    g = G.load("layernorm_c6")
    net = W.LayerNorm(6).cuda()
    net.load_state_dict(g["sd"])
    check(net(dev(g["inp"]["x"])), g["out"]["y"], what="layernorm")
    g = G.load("linearconv_k3_d2")
    net = W.LinearConv1d(4, 6, 3, dilation=2).cuda()
    net.load_state_dict(g["sd"])
    check(net.linear(dev(g["inp"]["frame"])), g["out"]["y"], what="linearconv.linear")
    # forward == linear applied at every valid frame
    x = torch.randn(2, 4, 21)
    yf = net(x.cuda()).cpu()
    ref = torch.nn.functional.conv1d(x, g["sd"]["weight"], g["sd"]["bias"], dilation=2)
    check(yf, ref, what="linearconv.forward")
    g = G.load("multiplicative_unit")
    net = W.MultiplicativeUnit(6, 3, dilation=2).cuda()
    net.load_state_dict(g["sd"])
    check(net(dev(g["inp"]["x"])), g["out"]["y"], what="mu")


def test_edge_shapes():
    """Ragged / tiny / tile-crossing shapes: T smaller than the dilation, T = 1, T straddling the 128-frame
    tile, channel counts that are not multiples of the tile."""
    torch.manual_seed(5)
    for (cin, cout, k, d, T, causal) in [(3, 5, 2, 8, 5, True), (3, 5, 2, 8, 5, False), (7, 130, 2, 3, 1, False),
                                         (70, 9, 3, 2, 129, False), (4, 4, 2, 1, 257, True),
                                         (65, 65, 2, 4, 300, False)]:
        blk = W.ResidualBlock(cin, cout, k, d, causal=causal)
        x = torch.randn(2, cin, T)
        sd = {kk: v.detach() for kk, v in blk.state_dict().items()}
        res_ref, skip_ref = O.residual_block(sd, "", x, d, causal)
        res, skip = blk.cuda()(x.cuda())
        check(res, res_ref, what="edge res %s" % ((cin, cout, k, d, T, causal),))
        check(skip, skip_ref, what="edge skip %s" % ((cin, cout, k, d, T, causal),))
    # empty batch
    blk = W.ResidualBlock(4, 4, 2, 1).cuda()
    res, skip = blk(torch.zeros(0, 4, 16).cuda())
    assert res.shape == (0, 4, 16)
    # non-contiguous view input, as legacy_code/train.py:30 passes sig[:, :, 0:-1]
    x = torch.randn(2, 4, 33)
    sd = {kk: v.detach().cpu() for kk, v in blk.state_dict().items()}
    res_ref, _ = O.residual_block(sd, "", x[:, :, 0:-1], 1, True)
    res, _ = blk(x.cuda()[:, :, 0:-1])
    check(res, res_ref, what="view input")


def _grads_vs_oracle(net, fwd_oracle, x, tol=2e-4):
    """Backward parity: gradients of sum(y * r) for a fixed random r, CUDA kernels vs torch autograd through
    the oracle on the same weights."""
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    y_ref = fwd_oracle(sd, xr)
    outs_ref = y_ref if isinstance(y_ref, tuple) else (y_ref,)
    rs = [torch.randn_like(o) for o in outs_ref]
    sum((o * r).sum() for o, r in zip(outs_ref, rs)).backward()
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    outs = y if isinstance(y, tuple) else (y,)
    sum((o * r.cuda()).sum() for o, r in zip(outs, rs)).backward()
    for o, orf in zip(outs, outs_ref):
        check(o, orf.detach(), what="fwd")
    assert G.rel_linf(xg.grad.cpu(), xr.grad) <= tol, ("dx", G.rel_linf(xg.grad.cpu(), xr.grad))
    for n, p in net.named_parameters():
        ref = sd[n].grad
        if ref is None:
            continue
        assert p.grad is not None, n
        e = G.rel_linf(p.grad.cpu(), ref)
        assert e <= tol, (n, e)


def test_backward_block():
    torch.manual_seed(11)
    for causal, k, d in [(True, 2, 2), (False, 2, 3), (False, 3, 2)]:
        blk = W.ResidualBlock(6, 10, k, d, causal=causal).cuda()
        x = torch.randn(3, 6, 50)
        _grads_vs_oracle(blk, lambda sd, xx: O.residual_block(sd, "", xx, d, causal), x)


def test_backward_networks():
    torch.manual_seed(12)
    layers = [(12, 12, 2, dd) for dd in (1, 2, 4)]
    net = W.WaveNet(12, 2, layers, 12, softmax=True).cuda()
    _grads_vs_oracle(net, lambda sd, xx: O.wavenet_forward(sd, xx, layers, softmax=True), torch.randn(2, 12, 40))
    net = W.RawCTCNet(12, 3, 5, layers, 12, softmax=False).cuda()
    _grads_vs_oracle(net, lambda sd, xx: O.raw_ctcnet_forward(sd, xx, layers, softmax=False), torch.randn(2, 1, 40))
    net = W.WaveNetClassifier(12, 5, layers, 12, pool_kernel_size=3, softmax=False).cuda()
    _grads_vs_oracle(net, lambda sd, xx: O.classifier_forward(sd, xx, layers, pool_kernel_size=3, softmax=False),
                     torch.randn(2, 12, 50))
    net = W.LayerNorm(7).cuda()
    _grads_vs_oracle(net, lambda sd, xx: O.layernorm(xx, sd["gamma"], sd["beta"]), torch.randn(2, 7, 19))
    net = W.MultiplicativeUnit(6, 3, dilation=2).cuda()        # fused four-conv launch + gate kernel, fwd and bwd
    _grads_vs_oracle(net, lambda sd, xx: O.multiplicative_unit(sd, "", xx, 2), torch.randn(2, 6, 33))


def test_xent_and_train_step_golden():
    """legacy_code/train.py:24-55 with the fused log-softmax+NLL kernel in place of the Python loop over t."""
    g = G.load("train_step_small")
    m = g["meta"]
    wl = [tuple(l) for l in m["wave_layers"]]
    cl = [tuple(l) for l in m["cls_layers"]]
    wn = W.WaveNet(m["dim"], 2, wl, m["dim"], softmax=False).cuda()
    cn = W.WaveNetClassifier(m["dim"], m["num_labels"], cl, m["dim"], pool_kernel_size=m["pool"], softmax=False).cuda()
    wn.load_state_dict(g["other"]["wsd"])
    cn.load_state_dict(g["other"]["csd"])
    sig = g["inp"]["sig"].cuda()
    B, _, T = sig.shape
    pred = wn(sig[:, :, 0:-1])
    trans = cn(pred)
    check(pred, g["out"]["pred"], what="pred")
    check(trans, g["out"]["trans"], what="trans")
    dense = W.ops.argmax_channels(sig[:, :, 1:].contiguous())
    assert torch.equal(dense.cpu(), torch.max(g["inp"]["sig"][:, :, 1:], dim=1)[1])
    xe = W.functional.cross_entropy_sum(pred, dense) / B
    assert abs(float(xe) - float(g["out"]["xe"])) <= 1e-5 * abs(float(g["out"]["xe"]))
    probs = trans.permute(2, 0, 1).contiguous()
    pl = torch.full((B,), probs.shape[0], dtype=torch.int32)
    ctc = torch.nn.functional.ctc_loss(torch.log_softmax(probs.float(), 2), g["inp"]["seq"].cuda(), pl,
                                       g["inp"]["lengths"], blank=0, reduction="sum")
    joint = xe / T + ctc / trans.shape[2]
    assert abs(float(joint) - float(g["out"]["joint"])) <= 1e-5 * abs(float(g["out"]["joint"]))
    joint.backward()
    for k, ref in g["other"]["wgrad"].items():
        e = G.rel_linf(dict(wn.named_parameters())[k].grad.cpu(), ref)
        assert e <= 2e-4, (k, e)
    for k, ref in g["other"]["cgrad"].items():
        e = G.rel_linf(dict(cn.named_parameters())[k].grad.cpu(), ref)
        assert e <= 2e-4, (k, e)


def test_bf16_generic_path():
    """bf16 storage through the generic kernels: compare with the oracle evaluated on the SAME bf16-rounded
    weights and input (SURVEY 7 hard part 1) -- <= 2e-2 relative on logits."""
    torch.manual_seed(21)
    layers = [(32, 32, 2, d) for d in (1, 2, 4, 8)]
    net = W.WaveNet(32, 2, layers, 32, softmax=False)
    sd = {k: v.detach().bfloat16().float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 32, 200).bfloat16()
    ref = O.wavenet_forward(sd, x.float(), layers, softmax=False)
    with torch.no_grad():
        y = net.cuda().bfloat16()(x.cuda())
    assert y.dtype == torch.bfloat16
    check(y, ref, tol=2e-2, what="bf16 wavenet")


def test_time_sharding_matches_full_read():
    """Config 5 in miniature: a read split into 4 time shards with receptive-field halos (ranks emulated one after
    the other on one GPU) gives the same logits as one pass, on both kernel paths."""
    from wavenet_speech_b200 import sharding as S
    torch.manual_seed(31)
    # fp32 generic path, odd sizes
    layers = [(12, 12, 2, d) for d in (1, 2, 4, 8)] + [(12, 12, 3, 3)]
    net = W.RawCTCNet(12, 3, 5, layers, 12, softmax=False).cuda()
    T = 997
    x = torch.randn(2, 1, T).cuda()
    hl, hr = S.raw_ctcnet_halo(net)
    with torch.no_grad():
        full = net(x)
        outs = []
        for r in range(4):
            plan = S.time_shard_plan(T, r, 4, hl, hr)
            outs.append(S.time_sharded_forward(net, x[:, :, plan["lo"]:plan["hi"]].contiguous(), plan, T,
                                               out_extra=net.feature_kwidth - 1))
    y = torch.cat(outs, 2)
    assert y.shape == full.shape
    assert G.rel_linf(y.cpu(), full.cpu()) <= 1e-6
    # bf16 tensor-core path
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8, 16)]
    net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False).cuda().bfloat16()
    T = 5000
    x = torch.randn(1, 1, T).cuda().bfloat16()
    hl, hr = S.raw_ctcnet_halo(net)
    with torch.no_grad():
        full = net(x)
        outs = []
        for r in range(3):
            plan = S.time_shard_plan(T, r, 3, hl, hr)
            outs.append(S.time_sharded_forward(net, x[:, :, plan["lo"]:plan["hi"]].contiguous(), plan, T,
                                               out_extra=net.feature_kwidth - 1))
    y = torch.cat(outs, 2)
    assert torch.equal(y, full)        # same frames, same arithmetic: bit-identical


def test_standalone_gate():
    torch.manual_seed(2)
    a = torch.randn(2, 5, 33, requires_grad=True)
    b = torch.randn(2, 5, 33, requires_grad=True)
    ref = O.gated_activation(a, b)
    r = torch.randn_like(ref)
    (ref * r).sum().backward()
    ag, bg = a.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    y = W.GatedActivationUnit()(ag, bg)
    (y * r.cuda()).sum().backward()
    check(y, ref.detach(), what="gate")
    assert G.rel_linf(ag.grad.cpu(), a.grad) <= 1e-5 and G.rel_linf(bg.grad.cpu(), b.grad) <= 1e-5


def test_positions_parameters_get_gradients():
    """RawCTCNet(positions=True) is trained by the reference (pretrain_tnt.py:121-124, tests/kmer_stay_prediction.py:52):
    the position layer's weight and bias receive the oracle's gradients (round 1 returned None for them)."""
    torch.manual_seed(21)
    layers = [(12, 12, 2, dd) for dd in (1, 2)]
    net = W.RawCTCNet(12, 3, 5, layers, 12, positions=True, softmax=False).cuda()
    with torch.no_grad():                     # spread w*t+b over (-1, 1) so that hardtanh's linear region is populated
        pc = net.positions_conv1x1[0]
        pc.weight.copy_((torch.rand_like(pc.weight) - 0.5) * 0.08)
        pc.bias.copy_((torch.rand_like(pc.bias) - 0.5) * 1.5)
    _grads_vs_oracle(net, lambda sd, xx: O.raw_ctcnet_forward(sd, xx, layers, positions=True, softmax=False),
                     torch.randn(2, 1, 40))
    assert net.positions_conv1x1[0].weight.grad is not None and net.positions_conv1x1[0].bias.grad is not None
    assert float(net.positions_conv1x1[0].weight.grad.abs().sum()) > 0


@pytest.mark.parametrize("k,d,causal", [(8, 1, True), (6, 3, False), (16, 2, False)])
def test_wide_kernel_block(k, d, causal):
    """Kernel widths above the 4 taps one launch takes (the reference trains with widths up to 32,
    pretrain_tnt.py:98,120): chained tap chunks + the stand-alone gate kernel, forward and backward."""
    torch.manual_seed(31 + k)
    blk = W.ResidualBlock(6, 10, k, d, causal=causal).cuda()
    _grads_vs_oracle(blk, lambda sd, xx: O.residual_block(sd, "", xx, d, causal), torch.randn(2, 6, 70))


def test_wide_featuriser():
    """RawCTCNet with feature_kwidth = 9 on the generic path (LeakyReLU applied after the chained tap chunks)."""
    torch.manual_seed(41)
    layers = [(8, 8, 2, 1)]
    net = W.RawCTCNet(8, 9, 5, layers, 8, softmax=False).cuda()
    _grads_vs_oracle(net, lambda sd, xx: O.raw_ctcnet_forward(sd, xx, layers, softmax=False), torch.randn(2, 1, 60))
