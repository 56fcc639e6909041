"""Tensor-core path (tcgen05 + TMA, bf16, NLC) against the CPU oracle evaluated on the SAME bf16-rounded
weights and inputs.  Tolerance from BASELINE.json's north_star: <= 2e-2 relative on logits for bf16."""
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


def test_layout_roundtrip():
    x = torch.randn(3, 70, 45).cuda()
    y = FP.ncl_to_nlc_bf16(x)
    assert torch.equal(y.float().cpu(), x.bfloat16().float().cpu().permute(0, 2, 1))
    z = FP.nlc_to_ncl(y, torch.float32)
    assert torch.equal(z.cpu(), x.bfloat16().float().cpu())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,C,T,pool", [(2, 64, 300, 3), (3, 256, 1001, 3), (1, 128, 16383, 3), (2, 70, 257, 2),
                                        (1, 64, 64, 1), (2, 64, 999, 7), (1, 64, 5, 3), (2, 64, 100, 40),
                                        (1, 33, 50, 3)])
def test_avgpool_to_nlc(B, C, T, pool, dt):
    """AvgPool1d(pool) fused with the NCL -> NLC bf16 layout change (classifier.py:53,102) on rows of any alignment,
    including the very last bytes of the tensor, partial channel / frame tiles and the fallback shapes."""
    from wavenet_speech_b200 import _lib, ops
    torch.manual_seed(B * 1000 + T)
    x = torch.randn(B, C, T).to(dt)
    To = T // pool
    ref = torch.nn.functional.avg_pool1d(x.float(), pool).bfloat16().float().permute(0, 2, 1) if To > 0 else None
    xg = x.cuda()
    y = torch.full((B, To, C), 7.0, dtype=torch.bfloat16, device="cuda")
    _lib.call("wnb200_avgpool_ncl_to_nlc", 1 if dt == torch.bfloat16 else 0, B, C, T, pool, ops._p(xg), _lib.ACT_BF16,
              ops._p(y), ops._stream())
    torch.cuda.synchronize()
    if To > 0:
        # fp32 accumulation order inside a window may differ from torch's: one bf16 ulp
        assert (y.float().cpu() - ref).abs().max() <= 2 ** -7 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,C,T,pool", [(2, 256, 1001, 3), (1, 128, 16383, 3), (2, 64, 130, 2), (1, 72, 64, 1),
                                        (2, 64, 999, 7), (1, 64, 2, 3), (3, 128, 200, 65)])
def test_avgpool_bwd_from_nlc(B, C, T, pool, dt):
    """Backward of AvgPool1d (classifier.py:102) fused with the NLC -> NCL layout change, against autograd."""
    from wavenet_speech_b200 import _lib, ops
    torch.manual_seed(T + pool)
    To = T // pool
    dh = torch.randn(B, To, C).bfloat16()
    x = torch.zeros(B, C, T, requires_grad=True)
    if To > 0:
        torch.nn.functional.avg_pool1d(x, pool).backward(dh.float().permute(0, 2, 1))
        ref = x.grad
    else:
        ref = torch.zeros(B, C, T)
    dx = torch.full((B, C, T), 7.0, dtype=dt, device="cuda")
    _lib.call("wnb200_avgpool_bwd_nlc_to_ncl", 1 if dt == torch.bfloat16 else 0, B, C, T, pool, ops._p(dh.cuda()),
              ops._p(dx), ops._stream())
    torch.cuda.synchronize()
    want = ref.to(dt).float() if dt == torch.bfloat16 else ref
    assert torch.equal(dx.float().cpu(), want)


@pytest.mark.parametrize("C,k,T,B", [(64, 2, 128, 1), (64, 1, 200, 2), (128, 2, 300, 2), (256, 2, 257, 2),
                                     (256, 3, 100, 1)])
def test_single_contraction(C, k, T, B):
    """stage 1 only: causal conv C->C with kernel k (the WaveNet entry conv, wavenet.py:54,93)."""
    torch.manual_seed(C + k + T)
    conv = W.CausalConv1d(C, C, k, dilation=1)
    with torch.no_grad():
        conv.conv1d.bias.add_(torch.randn(C) * 0.1)
    x = r16(torch.randn(B, C, T))
    ref = O.causal_conv1d(x, r16(conv.conv1d.weight), conv.conv1d.bias.detach(), 1)
    xn = FP.ncl_to_nlc_bf16(x.cuda())
    y = torch.empty_like(xn)
    FP.chain(xn, C, conv.offsets, FP.TC_LINEAR, FP._bf16(FP._taps_matrix(conv.conv1d.weight)).cuda(),
             conv.conv1d.bias.detach().float().cuda(), C, y_nlc=y)
    torch.cuda.synchronize()
    e = rel(y.float().permute(0, 2, 1), ref)
    assert e <= 1e-2, e


@pytest.mark.parametrize("Cin,N,k,T,B,leaky", [(64, 64, 2, 300, 2, 0), (256, 256, 2, 257, 3, 1), (128, 192, 1, 520, 1, 1),
                                               (256, 64, 3, 100, 2, 0)])
def test_dense_nlc(Cin, N, k, T, B, leaky):
    """CTA-pair dense contraction, NLC output (entry conv / first 1x1 of the head)."""
    torch.manual_seed(Cin + N + k)
    w = r16(torch.randn(N, Cin, k) / (Cin * k) ** 0.5)
    bias = torch.randn(N) * 0.1
    x = r16(torch.randn(B, Cin, T))
    offs = [j - (k - 1) for j in range(k)]
    ref = O.causal_conv1d(x, w, bias, 1)
    if leaky:
        ref = torch.nn.functional.leaky_relu(ref, 0.01)
    y = FP.dense(FP.ncl_to_nlc_bf16(x.cuda()), offs, FP._bf16(FP._taps_matrix(w)).cuda(), bias.cuda(), N, leaky=leaky)
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 2, 1), ref) <= 1e-2


@pytest.mark.parametrize("C,n_out,T,B,softmax,dt", [(256, 256, 384, 2, True, torch.bfloat16),
                                                    (256, 256, 130, 2, False, torch.bfloat16),
                                                    (128, 5, 200, 3, False, torch.float32),
                                                    (64, 8, 77, 2, True, torch.float32),
                                                    (256, 40, 257, 1, True, torch.bfloat16)])
def test_dense_head(C, n_out, T, B, softmax, dt):
    """CTA-pair dense contraction with the NCL / softmax epilogue (TMA-store and direct-store variants)."""
    torch.manual_seed(C + n_out + T)
    w = r16(torch.randn(n_out, C) / C ** 0.5)
    bias = torch.randn(n_out) * 0.1
    x = r16(torch.randn(B, C, T))
    ref = torch.nn.functional.conv1d(x, w.unsqueeze(2), bias)
    if softmax:
        ref = torch.softmax(ref, 1)
    n2 = (n_out + 15) // 16 * 16
    wp = torch.zeros(n2, C)
    wp[:n_out] = w
    bp = torch.zeros(n2)
    bp[:n_out] = bias
    out = torch.empty(B, n_out, T, dtype=dt, device="cuda")
    FP.dense(FP.ncl_to_nlc_bf16(x.cuda()), [0], FP._bf16(wp).cuda(), bp.cuda(), n2, mode=1, out=out, n_out=n_out,
             softmax=softmax)
    torch.cuda.synchronize()
    assert rel(out, ref) <= 1e-2, rel(out, ref)


@pytest.mark.parametrize("C,k,d,causal,T,B", [(64, 2, 1, True, 128, 1), (64, 2, 4, True, 300, 2),
                                              (128, 2, 2, False, 200, 2), (256, 2, 8, True, 384, 2),
                                              (256, 2, 3, False, 130, 3), (256, 3, 2, False, 260, 1),
                                              (256, 2, 512, True, 700, 1), (256, 1, 1, True, 64, 2)])
def test_fused_block(C, k, d, causal, T, B):
    """ResidualBlock + bottleneck in one launch vs block.py:54-82 + wavenet.py:100."""
    torch.manual_seed(C + 7 * k + d)
    blk = W.ResidualBlock(C, C, k, d, causal=causal)
    bn = torch.nn.Conv1d(C, C, 1)
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    sd = {kk: r16(v) if v.dim() > 1 else v.detach() for kk, v in blk.state_dict().items()}
    x = r16(torch.randn(B, C, T))
    res_ref, skip_ref = O.residual_block(sd, "", x, d, causal)
    # fold reference: bottleneck applied to the skip output, fp32 weights (the kernel rounds the folded product)
    contrib_ref = torch.nn.functional.conv1d(skip_ref, bn.weight.detach(), bn.bias.detach())
    pk = FP.pack_block(blk, bn)
    pk = {kk: (v.cuda() if torch.is_tensor(v) else v) for kk, v in pk.items()}
    xn = FP.ncl_to_nlc_bf16(x.cuda())
    res = torch.empty_like(xn)
    prev = torch.randn(B, T, C).cuda()
    skips = prev.clone()
    FP.chain(xn, C, pk["offsets"], FP.TC_GATE, pk["w1"], pk["b1"], 2 * C, n2=2 * C, use_x2=1,
             epi2=FP.EPI2_RESBLOCK, w2=pk["w2"], b2=pk["b2"], y_nlc=res, skips=skips, skips_init=0)
    torch.cuda.synchronize()
    e_res = rel(res.float().permute(0, 2, 1), res_ref)
    e_skip = rel((skips - prev).permute(0, 2, 1), contrib_ref)
    assert e_res <= BF16_TOL, ("res", e_res)
    assert e_skip <= BF16_TOL, ("skip", e_skip)
    # init mode writes instead of accumulating
    skips2 = torch.full_like(prev, 1e9)
    FP.chain(xn, C, pk["offsets"], FP.TC_GATE, pk["w1"], pk["b1"], 2 * C, n2=2 * C, use_x2=1,
             epi2=FP.EPI2_RESBLOCK, w2=pk["w2"], b2=pk["b2"], y_nlc=None, skips=skips2, skips_init=1)
    assert rel(skips2.permute(0, 2, 1), contrib_ref) <= BF16_TOL
    if C in (128, 256):
        # pipelined kernels (TMA store / reduce-add outputs): variant 1 = single CTA, 2 = CTA pair (cta_group::2)
        for variant in (1, 2):
            res3 = torch.empty_like(xn)
            skips3 = prev.clone()
            FP.resblock(xn, pk, res3, skips3, False, variant=variant)
            torch.cuda.synchronize()
            e = rel(res3.float().permute(0, 2, 1), res_ref)
            assert e <= BF16_TOL, ("res variant %d" % variant, e)
            assert rel((skips3 - prev).permute(0, 2, 1), contrib_ref) <= BF16_TOL, variant
            skips4 = torch.full_like(prev, 1e9)
            FP.resblock(xn, pk, None, skips4, True, variant=variant)
            assert rel(skips4.permute(0, 2, 1), contrib_ref) <= BF16_TOL, variant
            # the kernels agree closely with each other
            assert rel(res3.float(), res.float().cpu()) <= 1e-2, variant


@pytest.mark.parametrize("C,nl,T,B,softmax", [(64, 4, 300, 2, True), (128, 5, 1000, 2, False),
                                              (256, 6, 640, 2, True)])
def test_wavenet_tc(C, nl, T, B, softmax):
    torch.manual_seed(C + nl)
    layers = [(C, C, 2, 2 ** i) for i in range(nl)]
    net = W.WaveNet(C, 2, layers, C, softmax=softmax)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    lev = torch.randint(0, C, (B, T))
    x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0)
    ref = O.wavenet_forward(sd, x, layers, softmax=softmax)
    net = net.cuda().bfloat16()
    with torch.no_grad():
        net(x.cuda().bfloat16())                     # first call packs the weights on the device (cached afterwards)
        before = W._lib.launch_count
        y = net(x.cuda().bfloat16())
        assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(ref.shape)
    # transpose, entry, blocks (each stores its gate), the stack-wide skip contraction (C >= 128), 2 x head
    assert W._lib.launch_count - before == nl + (5 if C >= 128 else 4)
    e = rel(y, ref)
    assert e <= BF16_TOL, e
    if not softmax:
        agree = (y.float().cpu().argmax(1) == ref.argmax(1)).float().mean().item()
        assert agree > 0.97, agree


@pytest.mark.parametrize("C,nl,T,B,softmax,ldt", [(256, 3, 700, 2, True, torch.uint8), (128, 4, 257, 3, False, torch.int64),
                                                  (256, 2, 1, 1, True, torch.int32)])
def test_wavenet_forward_levels_equals_one_hot_call(C, nl, T, B, softmax, ldt):
    """forward_levels(levels) (entry conv as a gather of weight columns) is bit-identical to forward(one_hot(levels))."""
    torch.manual_seed(C + T)
    layers = [(C, C, 2, 2 ** i) for i in range(nl)]
    net = W.WaveNet(C, 2, layers, C, softmax=softmax)
    with torch.no_grad():
        net.entry_conv1d.conv1d.bias.add_(torch.randn(C) * 0.1)
    net = net.cuda().bfloat16().eval()
    lev = torch.randint(0, C, (B, T))
    x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0).cuda().bfloat16()
    with torch.no_grad():
        ref = net(x)
    y = net.forward_levels(lev.to(ldt).cuda())
    assert y.dtype == ref.dtype and torch.equal(y, ref)
    from wavenet_speech_b200.pipeline import HostPipeline
    lev_host = lev.to(ldt).pin_memory()
    out = HostPipeline(net, chunks=2, fn=net.forward_levels)(lev_host)
    assert torch.equal(out, ref.cpu())


@pytest.mark.parametrize("C,dil,T,B", [(256, [1, 2, 4], 700, 2), (128, [1, 2], 130, 3), (256, [4], 257, 1)])
def test_last_layer_emits_head_input_bitwise(C, dil, T, B):
    """The last block launch of an inference stack writes LeakyReLU(skip sum) as bf16 itself (TMA-loads the running
    sum, adds its tile in registers): bit-identical to reduce-add + the separate conversion pass, also for a one-layer
    stack (nothing to load) and for tiles that straddle T."""
    torch.manual_seed(C + T)
    net = W.WaveNet(C, 2, [(C, C, 2, d) for d in dil], C, softmax=False).cuda().bfloat16().eval()
    x = torch.randn(B, C, T, device="cuda").bfloat16()
    try:
        FP.DEFER_SKIP = False                     # this is about the in-HBM skip sum of the bf16 format (resblock2_kernel)
        with torch.no_grad(), FP.tc_precision("fast"):
            FP.FUSE_FINAL = False
            ref = net(x)
            FP.FUSE_FINAL = True
            y = net(x)
    finally:
        FP.FUSE_FINAL = True
        FP.DEFER_SKIP = True
    assert torch.equal(y, ref)


def test_time_slice_input_is_read_in_place():
    """train.py:30 feeds sig[:, :, 0:-1]: the layout kernel reads the non-contiguous slice in place (no copy), for
    inference and for the training forward / backward, with the same bits as a contiguous copy."""
    torch.manual_seed(3)
    C, T, B = 128, 520, 2
    net = W.WaveNet(C, 2, [(C, C, 2, d) for d in (1, 2, 4)], C, softmax=False).cuda()
    sig = torch.randn(B, C, T, device="cuda").bfloat16()
    view = sig[:, :, 0:-1]
    assert not view.is_contiguous()
    with torch.no_grad():
        assert torch.equal(net(view), net(view.contiguous()))
    outs = []
    for x in (view, view.contiguous()):
        net.zero_grad(set_to_none=True)
        y = net(x)
        y.float().square().mean().backward()
        outs.append((y.detach().clone(), net.entry_conv1d.conv1d.weight.grad.detach().clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    # (the weight-gradient kernels meet their per-CTA partial sums by fp32 reduce-add in no fixed order)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-3, atol=1e-7)
    assert torch.equal(FP.ncl_to_nlc_bf16(sig[:, 3:77, 5:400].float()), FP.ncl_to_nlc_bf16(sig[:, 3:77, 5:400].float().contiguous()))


@pytest.mark.parametrize("name", ["wavenet_c128_bf16w", "rawctcnet_c128_bf16w", "classifier_c128_bf16w"])
def test_tensor_core_path_against_reference_fixtures(name):
    """The tcgen05 path against outputs written by the REFERENCE's own modules (oracle/gen_golden_tc.py; weights and
    inputs bf16-representable, reference arithmetic fp32): logits within the stated bf16 tolerance, same per-frame
    argmax and the same collapsed greedy decode on (nearly) every frame."""
    g = G.load(name)
    m = g["meta"]
    if name.startswith("wavenet"):
        net = W.WaveNet(m["in_dim"], m["entry_kwidth"], m["layers"], m["out_dim"], softmax=m["softmax"])
    elif name.startswith("rawctcnet"):
        net = W.RawCTCNet(m["num_features"], m["feature_kwidth"], m["num_labels"], m["layers"], m["out_dim"],
                          positions=m["positions"], softmax=m["softmax"], causal=m["causal"])
    else:
        net = W.WaveNetClassifier(m["in_dim"], m["num_labels"], m["layers"], m["out_dim"],
                                  pool_kernel_size=m["pool_kernel_size"], softmax=m["softmax"])
    net.load_state_dict(g["sd"], strict=True)
    net = net.cuda().bfloat16().eval()
    ref = g["out"]["y"]
    with torch.no_grad():
        net(g["inp"]["x"].cuda().bfloat16())         # packs
        before = W._lib.launch_count
        y = net(g["inp"]["x"].cuda().bfloat16())
    assert W._lib.launch_count - before <= len(m["layers"]) + 6          # the fused (one launch per block) pipeline
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel(y, ref) <= BF16_TOL
    agree = (y.float().cpu().argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree >= 0.97, agree
    # fp32 tensors through the same kernels (reduced_precision): same tolerance, fp32 out
    net32 = net.float()
    with W.reduced_precision(True), torch.no_grad():
        y32 = net32(g["inp"]["x"].cuda())
    assert y32.dtype == torch.float32 and rel(y32, ref) <= BF16_TOL


def test_graphed_forward_replays_bitwise():
    """The whole forward captured in a CUDA graph (pipeline.GraphedForward) reproduces the eager result bit for bit,
    for new inputs as well, on the tensor-core and on the generic kernels."""
    from wavenet_speech_b200.pipeline import GraphedForward
    torch.manual_seed(11)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in (1, 2, 4, 8)], 256, softmax=False).cuda().bfloat16().eval()
    xs = [torch.randn(1, 1, 1000, device="cuda").bfloat16() for _ in range(3)]
    g = GraphedForward(net, xs[0])
    for x in xs:
        with torch.no_grad():
            ref = net(x)
        assert torch.equal(g(x), ref)
    net32 = W.WaveNet(16, 2, [(16, 16, 2, d) for d in (1, 2)], 16, softmax=True).cuda().eval()      # generic fp32 path
    x32 = torch.randn(2, 16, 100, device="cuda")
    g32 = GraphedForward(net32, x32)
    with torch.no_grad():
        assert torch.equal(g32(x32), net32(x32))


def test_reduced_precision_switch_routes_fp32_models():
    """An fp32 model with fp32 inputs stays on the fp32 FFMA kernels (<= 1e-5) unless `reduced_precision(True)`
    opts in to the tensor-core kernels: then outputs and gradients are fp32 tensors within the bf16 tolerance."""
    torch.manual_seed(5)
    C, T, B = 128, 700, 2
    layers = [(C, C, 2, 2 ** i) for i in range(4)]
    net = W.WaveNet(C, 2, layers, C, softmax=False)
    sd32 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd16 = {k: r16(v) for k, v in sd32.items()}
    lev = torch.randint(0, C, (B, T))
    x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0)
    ref32 = O.wavenet_forward(sd32, x, layers, softmax=False)
    ref16 = O.wavenet_forward(sd16, x, layers, softmax=False)
    net = net.cuda()
    xg = x.cuda()
    with torch.no_grad():
        y = net(xg)
    assert y.dtype == torch.float32 and rel(y, ref32) <= 1e-5            # default: true fp32 arithmetic
    with W.reduced_precision(True):
        with torch.no_grad():
            net(xg)                                  # packs
            before = W._lib.launch_count
            y2 = net(xg)
            assert W._lib.launch_count - before == len(layers) + 5         # the tensor-core launch sequence
        assert y2.dtype == torch.float32 and rel(y2, ref16) <= BF16_TOL
        xr = xg.clone().requires_grad_(True)
        out = net(xr)
        out.square().mean().backward()
        assert xr.grad is not None and xr.grad.dtype == torch.float32 and torch.isfinite(xr.grad).all()
        g = net.convolutions[1].conv_tanh.conv1d.weight.grad
        assert g is not None and g.dtype == torch.float32 and torch.isfinite(g).all() and float(g.abs().max()) > 0
    with torch.no_grad():
        y3 = net(xg)
    assert torch.equal(y3, y)                                               # the switch is scoped


@pytest.mark.parametrize("C,fk,T,B,causal,softmax", [(256, 3, 500, 2, False, False), (128, 1, 300, 3, True, True),
                                                     (256, 2, 4000, 2, False, False)])
def test_raw_ctcnet_tc(C, fk, T, B, causal, softmax):
    """RawCTCNet (ecoli-style: 256 ch, k=2, d in 1..16 x3 when C=256) on the tensor-core path."""
    torch.manual_seed(C + fk)
    dil = [1, 2, 4, 8, 16] * (3 if C == 256 else 1)
    layers = [(C, C, 2, d) for d in dil] + ([(C, C, 3, 2)] if C == 128 else [])
    net = W.RawCTCNet(C, fk, 5, layers, C, softmax=softmax, causal=causal)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    x = r16(torch.randn(B, 1, T))
    ref = O.raw_ctcnet_forward(sd, x, layers, softmax=softmax, causal=causal)
    net = net.cuda().bfloat16()
    with torch.no_grad():
        y = net(x.cuda().bfloat16())
    assert tuple(y.shape) == (B, 5, T + fk - 1) and y.dtype == torch.bfloat16
    e = rel(y, ref)
    # Deep untrained stacks amplify operand rounding ~1.45x per layer (hard part 1).  The default `precise`
    # format (fp16 operands, (hi, lo) stream, exact gate) holds 2e-2 at this depth with no envelope; the bf16 `fast`
    # format is held to "no worse than torch's own bf16 evaluation of the same network" (15 blocks: 7e-2 vs 1e-1).
    assert e <= BF16_TOL, e
    agree = (y.float().cpu().argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree >= 0.99, agree
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    t16 = O.raw_ctcnet_forward(sd16, x.bfloat16(), layers, softmax=softmax, causal=causal).float()
    e_torch = rel(t16, ref)
    with FP.tc_precision("fast"), torch.no_grad():
        yf = net(x.cuda().bfloat16())
    ef = rel(yf, ref)
    assert ef <= max(BF16_TOL, e_torch), (ef, e_torch)
    agree_f = (yf.float().cpu().argmax(1) == ref.argmax(1)).float().mean().item()
    agree_torch = (t16.argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree_f >= min(0.97, agree_torch - 0.01), (agree_f, agree_torch)


def test_classifier_tc():
    torch.manual_seed(9)
    C = 256
    layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16]]
    net = W.WaveNetClassifier(C, 5, layers, C, pool_kernel_size=3, softmax=False)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    x = r16(torch.randn(2, C, 1000))
    ref = O.classifier_forward(sd, x, layers, pool_kernel_size=3, softmax=False)
    with torch.no_grad():
        y = net.cuda().bfloat16()(x.cuda().bfloat16())
    assert tuple(y.shape) == (2, 5, 333)
    assert rel(y, ref) <= BF16_TOL, rel(y, ref)


def test_host_pipeline_matches_direct_call():
    """Chunked H2D / compute / D2H overlap returns exactly what the plain module call returns."""
    from wavenet_speech_b200.pipeline import HostPipeline
    torch.manual_seed(4)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4)]
    net = W.WaveNet(C, 2, layers, C, softmax=True).cuda().bfloat16()
    x = torch.randn(7, C, 500).bfloat16().pin_memory()
    with torch.no_grad():
        direct = net(x.cuda()).cpu()
    for chunks in (1, 3, 7, 16):
        y = HostPipeline(net, chunks=chunks)(x)
        assert torch.equal(y, direct), chunks


@pytest.mark.parametrize("B,T,Cg,m0,N,off", [(1, 64, 256, 0, 256, 0), (2, 300, 256, 0, 256, -3), (3, 1000, 512, 256, 256, 8),
                                             (2, 130, 256, 0, 128, 0), (4, 5000, 256, 0, 256, -512)])
def test_wgrad_tc(B, T, Cg, m0, N, off):
    """Time-contraction weight gradient with MN-major tensor-core operands vs an fp32 einsum."""
    torch.manual_seed(B + T + off)
    g = r16(torch.randn(B, T, Cg) * 0.1)
    x = r16(torch.randn(B, T, N))
    xs = torch.zeros_like(x)
    if off >= 0:
        xs[:, :T - off] = x[:, off:] if off < T else 0
    else:
        xs[:, -off:] = x[:, :T + off] if -off < T else 0
    ref = torch.einsum("btm,btn->mn", g[:, :, m0:m0 + 256].double(), xs.double()).float()
    dw = FP.wgrad(g.cuda().bfloat16(), x.cuda().bfloat16(), off=off, m0=m0)
    torch.cuda.synchronize()
    e = rel(dw, ref)
    assert e <= 2e-3, e
    # accumulates into an existing buffer
    dw2 = FP.wgrad(g.cuda().bfloat16(), x.cuda().bfloat16(), off=off, m0=m0, dw=dw.clone())
    assert rel(dw2, 2 * ref) <= 2e-3


@pytest.mark.parametrize("B,T,Cg,m0,N,offs", [(2, 300, 512, 256, 256, (-4, 0)), (1, 1000, 256, 0, 128, (0, 7)),
                                              (3, 777, 256, 0, 256, (-512, 1))])
def test_wgrad2_tc(B, T, Cg, m0, N, offs):
    """Two X tensors sharing the G tiles (both taps of a conv / [gate | x]): dw = [dw_0 | dw_1]."""
    torch.manual_seed(B + T)
    g = r16(torch.randn(B, T, Cg) * 0.1)
    xs = [r16(torch.randn(B, T, N)), r16(torch.randn(B, T, N))]

    def shifted(x, off):
        y = torch.zeros_like(x)
        if off >= 0:
            if off < T:
                y[:, :T - off] = x[:, off:]
        elif -off < T:
            y[:, -off:] = x[:, :T + off]
        return y

    ref = torch.cat([torch.einsum("btm,btn->mn", g[:, :, m0:m0 + 256].double(), shifted(x, o).double())
                     for x, o in zip(xs, offs)], 1).float()
    if ref.shape[0] < 256:
        ref = torch.cat([ref, torch.zeros(256 - ref.shape[0], ref.shape[1])], 0)
    dw = FP.wgrad2(g.cuda().bfloat16(), [x.cuda().bfloat16() for x in xs], offs, m0=m0)
    torch.cuda.synchronize()
    assert rel(dw, ref) <= 2e-3, rel(dw, ref)
