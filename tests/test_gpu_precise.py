"""The `precise` activation format of the tensor-core path (WNB200_ACT_F16X2: fp16 operands, residual stream carried
as an fp16 (hi, lo) pair, exact gate) against the oracle -- the north star's bf16-class tolerance (2e-2 on the logits,
fp32 oracle on the same bf16-rounded weights and inputs) WITHOUT an envelope at the depth of the benchmarked stacks
(config 2: 20 blocks; config 4: 1 + 15 blocks), per-frame argmax agreement >= 99 %, and every frame whose argmax differs
from the oracle's is a near-tie of the oracle's own logits."""
import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP
from oracle import wavenet_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
TOL = 2e-2          # north star: <= 2e-2 for bf16 on logits


def r16(t):
    return t.detach().bfloat16().float()


def _decode_agreement(y, ref, tol_abs):
    """(fraction of frames with the oracle's argmax, True if every disagreeing frame is a near-tie of the oracle:
    its top-2 margin is below 2 * tol_abs, i.e. the two evaluations are within the stated tolerance of each other)."""
    a, b = y.argmax(1), ref.argmax(1)
    agree = float((a == b).float().mean())
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    bad = a != b
    ties_only = bool((margin[bad] <= 2 * tol_abs).all()) if bad.any() else True
    return agree, ties_only


@pytest.mark.parametrize("C,k,d,causal,T,B", [(256, 2, 8, True, 384, 2), (128, 2, 2, False, 200, 2), (256, 3, 2, False, 260, 1),
                                              (256, 2, 512, True, 700, 1), (128, 1, 1, True, 64, 2)])
def test_precise_block(C, k, d, causal, T, B):
    """One fused block in the fp16 (hi, lo) format: hi + lo reproduces the fp32 residual output to ~1e-3 (the gate is
    fp16), ten times inside the bf16 format's per-block error."""
    torch.manual_seed(C + 7 * k + d)
    blk = W.ResidualBlock(C, C, k, d, causal=causal)
    bn = torch.nn.Conv1d(C, C, 1)
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    sd = {kk: r16(v) if v.dim() > 1 else v.detach() for kk, v in blk.state_dict().items()}
    x = torch.randn(B, C, T) * 3.0
    x_hi = x.half().float()
    x_lo = (x - x_hi).half().float()
    res_ref, skip_ref = O.residual_block(sd, "", x_hi + x_lo, d, causal)
    # the dilated taps read the hi half only: the oracle's gate on x_hi, its projection on x_hi + x_lo
    res_hi_only, skip_hi = O.residual_block(sd, "", x_hi, d, causal)
    proj_lo = torch.nn.functional.conv1d(x_lo, sd["residual_proj.weight"].unsqueeze(2))
    res_model = res_hi_only + proj_lo
    contrib_ref = torch.nn.functional.conv1d(skip_hi, bn.weight.detach(), bn.bias.detach())
    pk = FP.pack_block(blk, bn, precise=True)
    pk = {kk: (v.cuda() if torch.is_tensor(v) else v) for kk, v in pk.items()}
    assert pk["w1h"].dtype == torch.float16 and pk["fmt"] == W._lib.ACT_F16X2
    to_nlc = lambda t: t.permute(0, 2, 1).contiguous().half().cuda()
    xh, xl = to_nlc(x_hi), to_nlc(x_lo)
    res, res_lo = torch.empty_like(xh), torch.empty_like(xh)
    prev = torch.randn(B, T, C).cuda()
    skips = prev.clone()
    FP.resblock(xh, pk, res, skips, False, x_lo=xl, res_lo=res_lo)
    torch.cuda.synchronize()
    got = (res.float() + res_lo.float()).permute(0, 2, 1).cpu()
    assert G.rel_linf(got, res_model) <= 2e-3, G.rel_linf(got, res_model)
    assert G.rel_linf(got, res_ref) <= 5e-3, G.rel_linf(got, res_ref)       # + the taps' fp16 rounding of x
    assert G.rel_linf((skips - prev).permute(0, 2, 1).cpu(), contrib_ref) <= 5e-3
    # hi is fp16(v), lo the remainder: |lo| <= half an ulp of hi
    assert float((res_lo.float().abs() - res.float().abs() * 2.0 ** -10).max()) <= 2.0 ** -24 + 1e-12
    # no lo input (first block after an exact fp16 producer): same as lo = 0
    res2, res2_lo = torch.empty_like(xh), torch.empty_like(xh)
    FP.resblock(xh, pk, res2, prev.clone(), False, x_lo=None, res_lo=res2_lo)
    res3, res3_lo = torch.empty_like(xh), torch.empty_like(xh)
    FP.resblock(xh, pk, res3, prev.clone(), False, x_lo=torch.zeros_like(xh), res_lo=res3_lo)
    assert torch.equal(res2, res3) and torch.equal(res2_lo, res3_lo)
    got2 = (res2.float() + res2_lo.float()).permute(0, 2, 1).cpu()
    assert G.rel_linf(got2, res_hi_only) <= 2e-3


def test_dense_split_output():
    """The producers of a residual stream (entry conv, RawCTCNet feature 1x1) write the fp16 (hi, lo) pair."""
    torch.manual_seed(5)
    B, T, Cin, N = 2, 300, 128, 256
    x = torch.randn(B, T, Cin).half()
    w = (torch.randn(N, 2 * Cin) * 0.2).half()
    bias = torch.randn(N)
    hi, lo = FP.dense(x.cuda(), [-1, 0], w.cuda(), bias.cuda(), N, leaky=1, fmt=W._lib.ACT_F16X2, split=True)
    xs = torch.cat([torch.nn.functional.pad(x.float(), (0, 0, 1, 0))[:, :T], x.float()], 2)
    ref = torch.nn.functional.leaky_relu(xs @ w.float().t() + bias, 0.01)
    got = hi.float().cpu() + lo.float().cpu()
    assert G.rel_linf(got, ref) <= 1e-5
    assert torch.equal(hi.cpu(), got.half()) or G.rel_linf(hi.float().cpu(), ref) <= 2.0 ** -10
    single = FP.dense(x.cuda(), [-1, 0], w.cuda(), bias.cuda(), N, leaky=1, fmt=W._lib.ACT_F16X2)
    assert torch.equal(single, hi)


DIL2 = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2


@pytest.mark.parametrize("seed", [0, 1])
def test_wavenet_config2_depth_meets_tolerance(seed):
    """Config 2's network (2 x (1..512), 256 channels, 20 blocks) on a short batch the oracle evaluates in seconds:
    logits and softmax within 2e-2 of the fp32 oracle, no envelope."""
    torch.manual_seed(seed)
    C = 256
    layers = [(C, C, 2, d) for d in DIL2]
    B, T = 2, 1500
    lev = torch.randint(0, C, (B, T))
    x = torch.zeros(B, C, T).scatter_(1, lev.unsqueeze(1), 1.0)
    for softmax in (False, True):
        net = W.WaveNet(C, 2, layers, C, softmax=softmax)
        sd = {k: r16(v) for k, v in net.state_dict().items()}
        ref = O.wavenet_forward(sd, x, layers, softmax=softmax)
        net = net.cuda().bfloat16().eval()
        with torch.no_grad():
            y = net(x.cuda().bfloat16()).float().cpu()
        err = G.rel_linf(y, ref)
        assert err <= TOL, (softmax, err)
        # 256 logits of an untrained net on uniformly random levels sit close together and the module's output dtype is
        # bf16: ~1 % of the frames are ties of the top two after the OUTPUT rounding alone (a one-block net shows the same
        # 98.9 %, profiles/r2_parity_vs_depth.json); on the benchmark's pore-model signal the agreement is 99.4 %
        # (tests/test_gpu_full_size.py holds it to >= 99 %).  Every disagreeing frame must be such a tie.
        agree, ties_only = _decode_agreement(y, ref, TOL * float(ref.abs().max()))
        assert agree >= 0.98 or softmax, agree
        assert ties_only
        if not softmax:
            with FP.tc_precision("fast"), torch.no_grad():
                yf = net(x.cuda().bfloat16()).float().cpu()
            assert err < G.rel_linf(yf, ref)          # and the bf16 format is measurably further away


@pytest.mark.parametrize("seed,causal", [(0, False), (1, False), (2, True)])
def test_raw_ctcnet_config4_depth_meets_tolerance(seed, causal):
    """Config 4's network (ecoli RawCTCNet: input block + 15 blocks, 256 channels): logits within 2e-2, per-frame argmax
    >= 99 %, the collapsed greedy decode differs only where the oracle's own logits tie within the tolerance."""
    torch.manual_seed(seed)
    C = 256
    layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3]
    net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False, causal=causal)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    from wavenet_speech_b200.utils import signal_gen as SG
    x = r16(torch.from_numpy(SG.raw_batch(3, 2000, seed=40 + seed)))
    ref = O.raw_ctcnet_forward(sd, x, layers, softmax=False, causal=causal)
    net = net.cuda().bfloat16().eval()
    with torch.no_grad():
        yb = net(x.cuda().bfloat16())
    y = yb.float().cpu()
    err = G.rel_linf(y, ref)
    assert err <= TOL, err
    agree, ties_only = _decode_agreement(y, ref, TOL * float(ref.abs().max()))
    assert agree >= 0.99, agree
    assert ties_only
    # device greedy decode of OUR logits == oracle decode of our logits (the decode kernel itself is exact) ...
    lab, n = W.ops.ctc_greedy_decode(yb)
    fr_ours = O.argmax_decode(y.permute(0, 2, 1))
    fr_ref = O.argmax_decode(ref.permute(0, 2, 1))
    same = 0
    for b in range(x.shape[0]):
        ours = [int(v) for v in O.collapse_decode(fr_ours[b])]
        assert lab[b, :int(n[b])].cpu().tolist() == ours
        same += ours == [int(v) for v in O.collapse_decode(fr_ref[b])]
    # ... and where every frame agrees the decoded sequences are identical by construction
    if agree == 1.0:
        assert same == x.shape[0]


def test_classifier_precise_depth():
    torch.manual_seed(3)
    C = 256
    layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3]
    net = W.WaveNetClassifier(C, 5, layers, C, pool_kernel_size=3, softmax=False)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    x = r16(torch.randn(2, C, 3000))
    ref = O.classifier_forward(sd, x, layers, pool_kernel_size=3, softmax=False)
    with torch.no_grad():
        y = net.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu()
    err = G.rel_linf(y, ref)
    assert err <= TOL, err


def test_fast_format_still_selectable_and_bitwise_stable():
    """tc_precision("fast") is the round-1 bf16 pipeline (and the training format); switching back and forth does not
    leak packs between the two formats."""
    torch.manual_seed(11)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8)]
    net = W.WaveNet(C, 2, layers, C, softmax=False).cuda().bfloat16().eval()
    x = torch.randn(2, C, 700, device="cuda").bfloat16()
    with torch.no_grad():
        yp = net(x)
        with FP.tc_precision("fast"):
            yf = net(x)
            yf2 = net(x)
        yp2 = net(x)
    assert torch.equal(yp, yp2) and torch.equal(yf, yf2)
    assert not torch.equal(yp, yf)


def test_host_pipeline_back_to_back_submits():
    """Two DIFFERENT batches submitted back to back: the second batch's copy-ins must wait for the first batch's last
    chunks, which still read the same two staging slots (the per-slot events were per-submit state)."""
    from wavenet_speech_b200.pipeline import HostPipeline
    torch.manual_seed(4)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8, 16, 32)]
    net = W.WaveNet(C, 2, layers, C, softmax=True).cuda().bfloat16().eval()
    xs = [torch.randn(8, C, 3000).bfloat16().pin_memory() for _ in range(3)]
    with torch.no_grad():
        direct = [net(x.cuda()).cpu() for x in xs]
    for chunks in (1, 2, 4):
        pipe = HostPipeline(net, chunks=chunks)
        outs = [pipe.submit(x) for x in xs]
        pipe.wait()
        for o, d in zip(outs, direct):
            assert torch.equal(o, d), chunks


def test_host_pipeline_graph_replay():
    """HostPipeline(graph=True): the chunk forward replayed as one CUDA graph per staging slot.  Different batches back
    to back (the graph's output buffer is reused: a replay waits for the previous copy-out), ragged last chunk (runs
    eagerly), levels in (forward_levels under capture): bit-identical to the direct calls."""
    from wavenet_speech_b200.pipeline import HostPipeline
    torch.manual_seed(6)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8, 16, 32)]
    net = W.WaveNet(C, 2, layers, C, softmax=True).cuda().bfloat16().eval()
    xs = [torch.randn(8, C, 3000).bfloat16().pin_memory() for _ in range(4)]
    with torch.no_grad():
        direct = [net(x.cuda()).cpu() for x in xs]
    for chunks in (1, 2, 4, 3):                     # 3: chunks of 2, 3, 3 reads -> the first one is ragged
        pipe = HostPipeline(net, chunks=chunks, graph=True)
        outs = [pipe.submit(x) for x in xs]
        pipe.wait()
        for o, d in zip(outs, direct):
            assert torch.equal(o, d), chunks
        outs = [pipe.submit(x) for x in reversed(xs)]            # the captured graphs again, other data
        pipe.wait()
        for o, d in zip(outs, reversed(direct)):
            assert torch.equal(o, d), chunks
    # new weights / another activation format after the capture: the pipeline captures again
    pipe = HostPipeline(net, chunks=2, graph=True)
    y_a = pipe(xs[0]).clone()
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)
        direct_b = net(xs[0].cuda()).cpu()
    y_b = pipe(xs[0]).clone()
    assert torch.equal(y_b, direct_b) and not torch.equal(y_a, y_b)
    with FP.tc_precision("fast"), torch.no_grad():
        direct_c = net(xs[0].cuda()).cpu()
        y_c = pipe(xs[0]).clone()
    assert torch.equal(y_c, direct_c)
    assert torch.equal(pipe(xs[0]), direct_b)
    levs = [torch.randint(0, C, (8, 3000), dtype=torch.uint8).pin_memory() for _ in range(3)]
    with torch.no_grad():
        direct = [net.forward_levels(l.cuda()).cpu() for l in levs]
    pipe = HostPipeline(net, chunks=2, fn=net.forward_levels, graph=True)
    outs = [pipe.submit(l) for l in levs]
    pipe.wait()
    for o, d in zip(outs, direct):
        assert torch.equal(o, d)


@pytest.mark.parametrize("fmt", ["fast", "precise"])
@pytest.mark.parametrize("L,B,T,C", [(3, 2, 300, 256), (5, 1, 130, 128), (20, 1, 257, 256)])
def test_stack_wide_skip_contraction(fmt, L, B, T, C):
    """wnb200_dense_fwd_tc with `nlayers`: y = LeakyReLU(sum_l W_l g_l + bias) over a gate stack [L, B, T, C] -- the skip
    sum of a whole residual stack as one contraction (K = L * C, accumulated in tensor memory)."""
    torch.manual_seed(L * 100 + C)
    dt = torch.float16 if fmt == "precise" else torch.bfloat16
    g = (torch.rand(L, B, T, C) * 2 - 1).to(dt)
    w = (torch.randn(C, L * C) / (L * C) ** 0.5).to(dt)
    bias = torch.randn(C) * 0.1
    y = FP.dense(g.cuda(), [0], w.cuda(), bias.cuda(), C, leaky=1,
                 fmt=W._lib.ACT_F16X2 if fmt == "precise" else W._lib.ACT_BF16, nlayers=L)
    torch.cuda.synchronize()
    ref = bias.view(1, 1, C) + sum(g[l].float() @ w[:, l * C:(l + 1) * C].float().t() for l in range(L))
    ref = torch.nn.functional.leaky_relu(ref, 0.01)
    assert y.dtype == dt and tuple(y.shape) == (B, T, C)
    assert G.rel_linf(y.float().cpu(), ref) <= (2e-3 if fmt == "precise" else 1e-2)


@pytest.mark.parametrize("fmt", ["fast", "precise"])
def test_deferred_skip_equals_running_sum_pipeline(fmt):
    """The two ways to the skip sum (gates stored + one stack-wide contraction / a running fp32 sum in HBM updated by
    every layer) agree to rounding: same gates, fp32 accumulation in both, only the summation order and the rounding of
    the folded weights' products differ."""
    torch.manual_seed(17)
    C = 256
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8, 16, 32)]
    net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False).cuda().bfloat16().eval()
    x = torch.randn(3, 1, 900, device="cuda").bfloat16()
    try:
        with torch.no_grad(), FP.tc_precision(fmt):
            y = net(x)
            FP.DEFER_SKIP = False
            y0 = net(x)
    finally:
        FP.DEFER_SKIP = True
    assert G.rel_linf(y.float().cpu(), y0.float().cpu()) <= 8e-3


def test_gate_stack_budget_chunks_the_batch():
    """A batch whose gate stack (layers x batch x frames x channels x 2 bytes) exceeds the budget is walked in batch
    chunks; one read that does not fit falls back to the in-HBM running sum.  Same bits either way."""
    torch.manual_seed(23)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4)]
    net = W.RawCTCNet(C, 3, 5, layers, C, softmax=False).cuda().bfloat16().eval()
    x = torch.randn(5, 1, 700, device="cuda").bfloat16()
    with torch.no_grad():
        ref = net(x)
        old = FP.GATE_STACK_BUDGET
        try:
            FP.GATE_STACK_BUDGET = 4 * 2 * 702 * C * 2 + 1            # room for two reads of the 4-block stack
            y = net(x)
            one = net(x[:1])                                          # a single read does not fit: running-sum pipeline
        finally:
            FP.GATE_STACK_BUDGET = old
    assert torch.equal(y, ref)
    assert G.rel_linf(one.float().cpu(), ref[:1].float().cpu()) <= 8e-3


def test_fp16_range_guard_falls_back_to_bf16_format():
    """A stream that leaves the fp16 range (here: the projections scaled so that it grows 30x per block) raises the block
    kernel's saturation flag; the first forward of that weight version notices, warns, and repeats in the bf16 format
    -- the result is bitwise the tc_precision("fast") one and finite -- and the model stays on that format until its
    weights change.  A well-scaled model never warns and is marked verified after one forward."""
    import warnings
    torch.manual_seed(5)
    C = 128
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8, 16, 32)]
    net = W.WaveNet(C, 2, layers, C, softmax=False).cuda().bfloat16().eval()
    x = torch.randn(2, C, 600, device="cuda").bfloat16()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)
        y_ok = net(x)
        assert FP.stream_saturated(net) is False
        saved = [b.residual_proj.weight.clone() for b in net.convolutions]
        for b in net.convolutions:
            b.residual_proj.weight.mul_(30.0)
    with torch.no_grad():
        with FP.tc_precision("fast"):
            y_fast = net(x)
        with pytest.warns(RuntimeWarning, match="fp16 range"):
            y = net(x)
        assert FP.stream_saturated(net) is True
        assert torch.isfinite(y.float()).all() and float(y.float().abs().max()) > 0
        assert torch.equal(y, y_fast)
        with warnings.catch_warnings():
            warnings.simplefilter("error", RuntimeWarning)
            assert torch.equal(net(x), y_fast)             # remembered: no second warning, no precise attempt
            for b, w in zip(net.convolutions, saved):
                b.residual_proj.weight.copy_(w)
            assert torch.equal(net(x), y_ok)               # new weight version: precise again
            assert FP.stream_saturated(net) is False


@pytest.mark.parametrize("C,mode", [(128, "precise"), (256, "precise"), (128, "fast")])
def test_raw_ctcnet_positions_on_the_tensor_core_path(C, mode):
    """RawCTCNet(positions=True) (raw_ctcnet.py:131-135): the position term hardtanh(w t + b) is added in the epilogue
    of the feature layer's 1x1 contraction, so the net stays on the tensor-core path (round 1 sent it to the generic
    kernels).  Against the oracle, and -- with a time offset t0 (a shard of a longer read) -- against the generic path."""
    torch.manual_seed(31 + C)
    layers = [(C, C, 2, d) for d in (1, 2, 4)]
    net = W.RawCTCNet(C, 3, 5, layers, C, positions=True, softmax=False)
    with torch.no_grad():                          # make the term vary over the read instead of saturating at t = 1
        net.positions_conv1x1[0].weight.mul_(1e-3).add_(torch.randn(C, 1, 1) * 2e-3)
        net.positions_conv1x1[0].bias.add_(torch.randn(C) * 0.3)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    from wavenet_speech_b200.utils import signal_gen as SG
    x = r16(torch.from_numpy(SG.raw_batch(2, 700, seed=5)))
    ref = O.raw_ctcnet_forward(sd, x, layers, positions=True, softmax=False)
    ref_nopos = O.raw_ctcnet_forward(sd, x, layers, positions=False, softmax=False)
    assert G.rel_linf(ref_nopos, ref) > 0.05       # the term matters in this test
    net = net.cuda().bfloat16().eval()
    xg = x.cuda().bfloat16()
    with FP.tc_precision(mode), torch.no_grad():
        net(xg)                                    # (weight packs are built on the first call)
        n0 = W._lib.launch_count
        y = net(xg).float().cpu()
        assert W._lib.launch_count - n0 <= 12      # featuriser, 1x1, 4 blocks, skip contraction, 2 head launches
        assert G.rel_linf(y, ref) <= TOL, G.rel_linf(y, ref)
        y_t0 = net(xg, t0=1234).float().cpu()
    # the generic path with the same offset (fp32 FFMA kernels on the bf16-rounded parameters)
    net32 = W.RawCTCNet(C, 3, 5, layers, C, positions=True, softmax=False)
    net32.load_state_dict(sd)
    net32 = net32.cuda().eval()
    with torch.no_grad():
        g_t0 = net32(x.cuda(), t0=1234).float().cpu()
    assert G.rel_linf(y_t0, g_t0) <= TOL, G.rel_linf(y_t0, g_t0)
    assert G.rel_linf(y_t0, y) > 1e-3


def test_hi_only_tail_of_a_stack():
    """The block kernel accepts `res` without `res_lo` (the stream's hi half only); fastpath.HI_ONLY_TAIL hands the last
    blocks of a stack over that way.  8 blocks with the last 4 hi-only: still within the tolerance, different bits from
    the all-pair run (the knob did something), and the default (0) is bit-identical to itself."""
    torch.manual_seed(5)
    C = 256
    layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16, 32, 64, 128]]
    net = W.WaveNet(C, 2, layers, C, softmax=False)
    sd = {k: r16(v) for k, v in net.state_dict().items()}
    lev = torch.randint(0, C, (2, 700))
    x = torch.zeros(2, C, 700).scatter_(1, lev.unsqueeze(1), 1.0)
    ref = O.wavenet_forward(sd, x, layers, softmax=False)
    net = net.cuda().bfloat16().eval()
    xg = x.cuda().bfloat16()
    assert FP.HI_ONLY_TAIL == 0
    with torch.no_grad():
        y0 = net(xg)
        try:
            FP.HI_ONLY_TAIL = 4
            y4 = net(xg)
            FP.HI_ONLY_TAIL = 100                    # more than the stack has: every block hi-only
            y_all = net(xg)
        finally:
            FP.HI_ONLY_TAIL = 0
        y0b = net(xg)
    assert torch.equal(y0, y0b)
    e0, e4, ea = (G.rel_linf(t.float().cpu(), ref) for t in (y0, y4, y_all))
    assert e0 <= TOL and e4 <= TOL and ea <= TOL, (e0, e4, ea)
    assert not torch.equal(y0, y_all)
