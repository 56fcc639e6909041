"""Secondary measurements (BASELINE configs 1, 3, 4, 5) -- not bench.py's headline line.  One JSON line each.
    python scripts/bench_extra.py [rawctc] [train] [longread] [example]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wavenet_speech_b200 as W
from wavenet_speech_b200 import sharding as S
from wavenet_speech_b200.utils import signal_gen as SG

which = set(a for a in sys.argv[1:] if "=" not in a) or {"rawctc", "train", "longread", "example"}
opts = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
ECOLI = [1, 2, 4, 8, 16] * 3


def timed(fn, steps, warmup=2):
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


def ecoli_net(fk=3):
    torch.manual_seed(0)
    return W.RawCTCNet(256, fk, 5, [(256, 256, 2, d) for d in ECOLI], 256, softmax=False)


if "rawctc" in which:      # config 4: RawCTCNet from configs/ecoli_testrun.json, batch sweep, bf16 inference
    net = ecoli_net().cuda().bfloat16().eval()
    flop = 17043456
    for B in (64, 256, 1024):
        x = torch.from_numpy(SG.raw_batch(min(B, 64), 4000, seed=7)).repeat((B + 63) // 64, 1, 1)[:B].cuda().bfloat16()
        with torch.no_grad():
            ms = timed(lambda: net(x), 5)
        sps = B * 4000 / (ms * 1e-3)
        print(json.dumps({"config": "rawctcnet_ecoli_fk3_bf16_fwd", "batch": B, "T": 4000, "ms_per_step": ms,
                          "samples_per_s": sps, "tflops_as_written": sps * flop / 1e12}))

if "graph" in which:       # small-batch basecalling latency: eager launches vs one CUDA-graph replay
    from wavenet_speech_b200.pipeline import GraphedForward
    net = ecoli_net().cuda().bfloat16().eval()
    for B in (1, 4, 16):
        x = torch.from_numpy(SG.raw_batch(B, 4000, seed=9)).cuda().bfloat16()
        with torch.no_grad():
            ms_eager = timed(lambda: net(x), 50, warmup=5)
        g = GraphedForward(net, x)
        ms_graph = timed(lambda: g(x), 50, warmup=5)
        print(json.dumps({"config": "rawctcnet_ecoli_fk3_bf16_fwd_latency", "batch": B, "T": 4000, "ms_eager": ms_eager,
                          "ms_cuda_graph": ms_graph, "samples_per_s_graph": B * 4000 / (ms_graph * 1e-3)}))

if "example" in which:     # config 1: RawCTCNet from configs/example.json, fp32, 8 x 4000 (generic fp32 kernels)
    torch.manual_seed(0)
    layers = [(1, 1, 1, 1)]
    net = W.RawCTCNet(256, 2, 8, layers, 256, softmax=False).cuda()
    xg = torch.from_numpy(SG.raw_batch(8, 4000, seed=3)).cuda()
    with torch.no_grad():
        ms = timed(lambda: net(xg), 10)
    # parity of this configuration against the oracle: tests/test_gpu_parity.py::test_raw_ctcnet (golden fixtures)
    print(json.dumps({"config": "rawctcnet_example_json_fp32_fwd", "batch": 8, "T": 4000, "ms_per_step": ms,
                      "samples_per_s": 32000 / (ms * 1e-3)}))

if "train" in which:       # config 3: WaveNet-CTC train step (legacy_code/train.py:24-61), fp32 generic kernels
    torch.manual_seed(0)
    dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2
    wn = W.WaveNet(256, 2, [(256, 256, 2, d) for d in dil], 256, softmax=False).cuda()
    cn = W.WaveNetClassifier(256, 5, [(256, 256, 2, d) for d in ECOLI], 256, pool_kernel_size=3, softmax=False).cuda()
    opt = torch.optim.Adam(list(wn.parameters()) + list(cn.parameters()), lr=1e-5)
    B, T = 4, 4096
    lev, labels = SG.quantized_batch(B, T, seed=5, with_labels=True)
    sig = torch.from_numpy(SG.one_hot(lev)).cuda()
    lengths = torch.tensor([min(len(l), 300) for l in labels], dtype=torch.int32)
    seq = torch.cat([torch.from_numpy(l[:300]) for l in labels]).int().cuda()

    def step():
        opt.zero_grad(set_to_none=True)
        pred = wn(sig[:, :, 0:-1])
        trans = cn(pred)
        dense = W.ops.argmax_channels(sig[:, :, 1:].contiguous())
        xe = W.functional.cross_entropy_sum(pred, dense) / B
        probs = trans.permute(2, 0, 1).contiguous()
        pl = torch.full((B,), probs.shape[0], dtype=torch.int32)
        ctc = torch.nn.functional.ctc_loss(torch.log_softmax(probs.float(), 2), seq, pl, lengths, blank=0,
                                           reduction="sum", zero_infinity=True)
        loss = xe / T + ctc / trans.shape[2]
        loss.backward()
        opt.step()
        return float(loss)

    l0 = step()
    ms = timed(step, 3, warmup=1)
    l1 = step()
    print(json.dumps({"config": "wavenet_ctc_train_step_fp32_generic", "batch": B, "T": T, "ms_per_step": ms,
                      "samples_per_s": B * T / (ms * 1e-3), "loss_first": l0, "loss_after": l1,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))

if "train_tc" in which:    # config 3 on the tensor-core path: bf16 operands, fp32 master weights / gradients / Adam
    from wavenet_speech_b200 import _lib
    torch.manual_seed(0)
    dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2
    wn = W.WaveNet(256, 2, [(256, 256, 2, d) for d in dil], 256, softmax=False).cuda()
    cn = W.WaveNetClassifier(256, 5, [(256, 256, 2, d) for d in ECOLI], 256, pool_kernel_size=3, softmax=False).cuda()
    if opts.get("opt", "ours") == "torch":
        opt = torch.optim.Adam(list(wn.parameters()) + list(cn.parameters()), lr=1e-5, fused=True)
    else:
        opt = W.optim.Adam(list(wn.parameters()) + list(cn.parameters()), lr=1e-5)
    B, T = int(opts.get("B", 32)), int(opts.get("T", 16384))
    lev, labels = SG.quantized_batch(min(B, 8), T, seed=5, with_labels=True)
    rep = (B + 7) // 8
    sig = torch.from_numpy(SG.one_hot(lev)).repeat(rep, 1, 1)[:B].cuda().bfloat16()
    labels = (labels * rep)[:B]
    nlab = T // 3 // 2
    lengths = torch.tensor([min(len(l), nlab) for l in labels], dtype=torch.int32)
    seq = torch.cat([torch.from_numpy(l[:nlab]) for l in labels]).int().cuda()

    def ctc_torch(trans):
        probs = trans.permute(2, 0, 1).contiguous()
        pl = torch.full((B,), probs.shape[0], dtype=torch.int32)
        return torch.nn.functional.ctc_loss(torch.log_softmax(probs.float(), 2), seq, pl, lengths, blank=0,
                                            reduction="sum", zero_infinity=True)

    def ctc_ours(trans):                    # device CTC kernels, reads the classifier's (B, C, T) output in place
        return W.functional.ctc_loss_sum(trans, seq, lengths, layout="bct")

    ctc_fn = ctc_torch if opts.get("ctc") == "torch" else ctc_ours

    def step():
        opt.zero_grad(set_to_none=True)
        pred = wn(sig[:, :, 0:-1])
        trans = cn(pred)
        dense = W.ops.argmax_channels(sig[:, :, 1:].contiguous())
        xe = W.functional.cross_entropy_sum(pred, dense) / B
        ctc = ctc_fn(trans)
        loss = xe / T + ctc / trans.shape[2]
        loss.backward()
        opt.step()
        return loss

    l0 = float(step())
    ms = timed(step, int(opts.get("steps", 5)), warmup=2)
    l1 = float(step())
    _lib.kernel_timing(True)
    step()
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1 in _lib.kernel_timing(False):
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    # host/device split of one step: wall clock with a sync after each phase
    def phase(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3
    ph = {}
    opt.zero_grad(set_to_none=True)
    pred, ph["wavenet_fwd"] = phase(lambda: wn(sig[:, :, 0:-1]))
    trans, ph["classifier_fwd"] = phase(lambda: cn(pred))
    def xe_():
        dense = W.ops.argmax_channels(sig[:, :, 1:].contiguous())
        return W.functional.cross_entropy_sum(pred, dense) / B
    xe, ph["xent"] = phase(xe_)
    ctc, ph["ctc_fwd"] = phase(lambda: ctc_fn(trans))
    loss = xe / T + ctc / trans.shape[2]
    _, ph["backward"] = phase(loss.backward)
    _, ph["adam"] = phase(opt.step)
    flop = 3 * (21495808 + 16910848 / 3)
    print(json.dumps({"config": "wavenet_ctc_train_step_bf16_tensor_core", "ctc": opts.get("ctc", "wnb200"), "batch": B, "T": T, "ms_per_step": ms,
                      "samples_per_s": B * T / (ms * 1e-3), "tflops_as_written_3x_fwd": B * T / (ms * 1e-3) * flop / 1e12,
                      "loss_first": l0, "loss_after": l1, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                      "phases_ms": {k: round(v, 2) for k, v in ph.items()},
                      "kernels_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}))

if "rawctc_train" in which:   # RawCTCNet CTC training step (legacy_code/run_raw_ctc.py:53-66) on the tensor-core path
    from wavenet_speech_b200 import _lib
    torch.manual_seed(0)
    net = ecoli_net().cuda()
    bn = torch.nn.BatchNorm1d(1).cuda()                     # run_raw_ctc.py:38,58 normalises the raw signal (outside the path)
    opt = torch.optim.Adam(list(net.parameters()) + list(bn.parameters()), lr=1e-5, weight_decay=1e-4, fused=True)
    B, T = int(opts.get("B", 128)), int(opts.get("T", 4000))
    sigs, labels = [], []
    rng = __import__("numpy").random.default_rng(9)
    for _ in range(min(B, 16)):
        sg_, lb_ = SG.raw_signal(T, rng, with_labels=True)
        sigs.append(torch.from_numpy(sg_).float())
        labels.append(torch.from_numpy(lb_))
    rep = (B + 15) // 16
    x = torch.stack(sigs).unsqueeze(1).repeat(rep, 1, 1)[:B].cuda()
    labels = (labels * rep)[:B]
    lengths = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    seq = torch.cat(labels).int().cuda()

    def step():
        opt.zero_grad(set_to_none=True)
        trans = net(bn(x).bfloat16())
        ctc = W.functional.ctc_loss_sum(trans, seq, lengths, layout="bct")
        loss = ctc / trans.shape[2]
        loss.backward()
        opt.step()
        return loss

    l0 = float(step())
    ms = timed(step, int(opts.get("steps", 5)), warmup=2)
    l1 = float(step())
    _lib.kernel_timing(True)
    step()
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1 in _lib.kernel_timing(False):
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    print(json.dumps({"config": "rawctcnet_ecoli_ctc_train_step_bf16_tensor_core", "batch": B, "T": T, "ms_per_step": ms,
                      "samples_per_s": B * T / (ms * 1e-3), "tflops_as_written_3x_fwd": B * T / (ms * 1e-3) * 3 * 17043456 / 1e12,
                      "loss_first": l0, "loss_after": l1, "labels_per_read": float(lengths.float().mean()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                      "kernels_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}))

if "longread" in which:    # config 5: 1M-sample read, time-sharded in 8 shards (emulated on one GPU) vs one pass
    net = ecoli_net().cuda().bfloat16().eval()
    T = 1000000
    x = torch.from_numpy(SG.raw_batch(1, T, seed=11)).cuda().bfloat16()
    with torch.no_grad():
        full = net(x)
        ms_full = timed(lambda: net(x), 3, warmup=1)
        hl, hr = S.raw_ctcnet_halo(net)
        world = 8
        outs, ms_shards = [], []
        for r in range(world):
            plan = S.time_shard_plan(T, r, world, hl, hr)
            x_ext = x[:, :, plan["lo"]:plan["hi"]].contiguous()
            outs.append(S.time_sharded_forward(net, x_ext, plan, T, out_extra=net.feature_kwidth - 1))
            ms_shards.append(timed(lambda: net(x_ext), 3, warmup=1))
        y = torch.cat(outs, 2)
    same = bool(torch.equal(y, full))
    print(json.dumps({"config": "rawctcnet_ecoli_1M_read_time_sharded", "T": T, "shards": world, "halo": [hl, hr],
                      "one_gpu_full_read_ms": ms_full, "one_gpu_samples_per_s": T / (ms_full * 1e-3),
                      "per_shard_ms_max": max(ms_shards), "projected_8gpu_samples_per_s": T / (max(ms_shards) * 1e-3),
                      "sharded_equals_full_bitwise": same,
                      "max_abs_diff": float((y.float() - full.float()).abs().max())}))


if "bytenet" in which:     # ByteNet residual blocks (block.py:86-173) and the frame-at-a-time LinearConv1d
    import torch.nn.functional as F
    sys.path.insert(0, ROOT)
    from oracle import wavenet_oracle as O          # stock-torch comparator on the same GPU (the reference's op sequence)
    C, k, d = int(opts.get("C", 512)), 3, 4
    B, T = int(opts.get("B", 16)), int(opts.get("T", 8192))
    for kind, cls, fn in (("relu", W.ResidualReLUBlock, O.residual_relu_block), ("mu", W.ResidualMUBlock, O.residual_mu_block)):
        for dtype in (torch.float32, torch.bfloat16):
            torch.manual_seed(0)
            net = cls(C, k, d)
            net.init()
            net = net.cuda().to(dtype)
            sd = {kk: v.detach() for kk, v in net.state_dict().items()}
            x = torch.randn(B, C, T, device="cuda", dtype=dtype)
            with torch.no_grad():
                ms = timed(lambda: net(x), 10)
                ms_stock = timed(lambda: fn(sd, "", x, d), 5)
                y, yr = net(x), fn(sd, "", x, d)
                ms_strict = None
                if dtype == torch.float32:      # stock torch defaults to TF32 convolutions: time it at fp32 accuracy too
                    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
                    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
                    ms_strict = timed(lambda: fn(sd, "", x, d), 5)
                    ys = fn(sd, "", x, d)
                    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
            h = C // 2
            # FLOP as written: 1x1 down + (relu: k-tap conv | mu: 4 k-tap + 4 1-tap convs) + 1x1 up, per frame
            mac = C * h + h * C + (k * h * h if kind == "relu" else 4 * k * h * h + 4 * h * h)
            xg = x.clone().requires_grad_(True)
            def step():
                for p in net.parameters():
                    p.grad = None
                net(xg).sum().backward()
            ms_tr = timed(step, 5)
            es = x.element_size()
            print(json.dumps({"what": "bytenet_block", "kind": kind, "dtype": str(dtype).split(".")[1], "C": C, "k": k, "d": d,
                              "B": B, "T": T, "fwd_ms": round(ms, 3), "fwd_samples_per_s": round(B * T / ms * 1e3),
                              "fwd_tflops": round(2 * mac * B * T / ms / 1e9, 1),
                              "min_bytes_gbps": round(2 * B * C * T * es / ms / 1e6, 1),
                              "stock_torch_gpu_fwd_ms": round(ms_stock, 3), "speedup_vs_stock": round(ms_stock / ms, 2),
                              "fwd_bwd_ms": round(ms_tr, 3),
                              "max_abs_diff_vs_stock": float((y.float() - yr.float()).abs().max()),
                              "stock_torch_gpu_strict_fp32_fwd_ms": None if ms_strict is None else round(ms_strict, 3),
                              "max_abs_diff_vs_strict_fp32_stock": None if ms_strict is None else
                              float((y.float() - ys.float()).abs().max())}))
    # incremental decoding: one frame per step through a LinearConv1d
    for (cin, cout, kk, dd, N) in ((512, 512, 3, 4, 16), (1024, 1024, 5, 16, 8)):
        conv = W.LinearConv1d(cin, cout, kk, dilation=dd).cuda()
        rf = conv.receptive_field
        st = conv.stream(N)
        xt = torch.randn(N, cin, device="cuda")
        frame = torch.randn(N, cin, rf, device="cuda")
        with torch.no_grad():
            us_push = timed(lambda: st.push(xt), 200, warmup=20) * 1e3
            us_lin = timed(lambda: conv.linear(frame), 200, warmup=20) * 1e3
            us_stock = timed(lambda: F.linear(frame[:, :, conv._ker_ixs].reshape(N, cin * kk),
                                              conv.weight.view(cout, cin * kk), conv.bias), 200, warmup=20) * 1e3
            g = torch.cuda.CUDAGraph()
            xs = xt.clone()
            yb = torch.empty(N, cout, device="cuda")
            from wavenet_speech_b200 import ops as _ops
            hist = torch.zeros(rf, N, cin, device="cuda")
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                _ops.linear_step(xs, conv.weight, conv.bias, dd, 0, hist, out=yb)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    for i in range(rf):          # one ring revolution: the slot addresses repeat with period rf
                        _ops.linear_step(xs, conv.weight, conv.bias, dd, i, hist, out=yb)
            us_graph = timed(lambda: g.replay(), 20, warmup=3) * 1e3 / rf
        wbytes = cout * cin * kk * 4
        print(json.dumps({"what": "linearconv_step", "cin": cin, "cout": cout, "k": kk, "d": dd, "rf": rf, "N": N,
                          "push_us": round(us_push, 2), "push_graph_us": round(us_graph, 2),
                          "weight_gbps_in_graph": round(wbytes / us_graph / 1e3, 1),
                          "linear_us": round(us_lin, 2), "stock_torch_linear_us": round(us_stock, 2)}))
