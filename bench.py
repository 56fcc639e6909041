#!/usr/bin/env python
"""Headline benchmark of the hot path: raw-signal samples/sec of the WaveNet forward pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): vanilla WaveNet,
2 stacks x 10 dilations (1..512), 256 residual/skip channels, 256-way softmax, batch 32 x 16384 samples
per GPU, bf16 forward.  Input: synthetic Gaussian pore-model signal (wavenet_speech_b200/utils/signal_gen.py),
mu-law quantised to 256 levels and one-hot encoded, exactly what the reference feeds its WaveNet.

One JSON line is printed by rank 0 (see the task contract): value = samples/s with inputs resident in HBM,
e2e = the same through the module call with pinned HOST buffers (H2D of the input and D2H of the output inside
the timed region), roofline = the fused residual-block kernel against the measured bf16 peak, cpu_baseline =
the oracle (a CPU restatement of the reference's PyTorch modules) on this box's host cores.

N > 1: one process per GPU (torchrun), the batch is sharded (each rank owns `batch` reads: weak scaling), no
data-path collective exists for inference; value = all ranks' samples / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (in_dim, entry_k, dilations, C, out_dim, batch, T)
    "wavenet_2x10_c256_b32_t16384": dict(in_dim=256, entry_k=2, dil=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2,
                                         C=256, batch=32, T=16384),
    "wavenet_small": dict(in_dim=64, entry_k=2, dil=[1, 2, 4, 8], C=64, batch=2, T=2048),
}
DEFAULT_WORKLOAD = "wavenet_2x10_c256_b32_t16384"
METRIC = "raw-signal samples/sec (WaveNet forward)"
UNIT = "samples/s"


def flops_per_timestep(w):
    C, D = w["C"], w["in_dim"]
    return 2 * w["entry_k"] * D * C + len(w["dil"]) * 16 * C * C + 4 * C * C


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class quiet_gc(object):
    """The cyclic garbage collector off for a timed region (after one explicit collection): a generation-2 pass over the
    process's objects takes tens of milliseconds of the host thread -- inside a 20-step region that starts from an
    empty launch queue it showed up as 1 run in 8 reading 40 M samples/s instead of 46 M."""

    def __enter__(self):
        import gc
        gc.collect()
        gc.disable()

    def __exit__(self, *a):
        import gc
        gc.enable()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.window = index, None, [], None

    def mark(self, t0, t1):
        self.window = (t0, t1)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up (NVML attach) can stall the driver for tens of milliseconds: wait for its first sample
            # so that this happens before the warm-up, never inside the timed region (measured: a 20-step region that
            # caught it read 36-38 M samples/s instead of 46 M)
            t_end = time.time() + 3.0
            while not self.lines and time.time() < t_end:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines
        if self.window is not None:
            inside = [l for l in lines if self.window[0] <= l[0] <= self.window[1] + 0.05]
            lines = inside if len(inside) >= 2 else lines
        for _ts, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_input(w, rank):
    """(B, in_dim, T) one-hot of a mu-law quantised synthetic pore-model signal, as a pinned host bf16 tensor."""
    from wavenet_speech_b200.utils import signal_gen as S
    lev = S.quantized_batch(w["batch"], w["T"], num_levels=w["in_dim"], seed=1234 + rank)
    lev = torch.from_numpy(lev)
    x = torch.zeros((w["batch"], w["in_dim"], w["T"]), dtype=torch.bfloat16)
    x.scatter_(1, lev.unsqueeze(1), 1.0)
    return x.pin_memory() if torch.cuda.is_available() else x


def build_model(w, softmax=True):
    import wavenet_speech_b200 as W
    torch.manual_seed(0)
    layers = [(w["C"], w["C"], 2, d) for d in w["dil"]]
    return W.WaveNet(w["in_dim"], w["entry_k"], layers, w["C"], softmax=softmax), layers


def cpu_baseline(w, target_s=12.0, threads=None):
    """Oracle (CPU restatement of the reference's PyTorch modules) on the host cores, bounded sample."""
    from oracle import wavenet_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    net, layers = build_model(w)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    D = w["in_dim"]

    def run(B, T):
        lev = torch.randint(0, D, (B, T))
        x = torch.zeros(B, D, T).scatter_(1, lev.unsqueeze(1), 1.0)
        t0 = time.perf_counter()
        with torch.no_grad():
            O.wavenet_forward(sd, x, layers, softmax=True)
        return time.perf_counter() - t0

    run(1, 1024)                                  # warm-up (thread pool, mkldnn primitives)
    probe_T = 2048
    dt = run(1, probe_T)
    rate = probe_T / dt
    T = int(min(w["T"], max(2048, rate * target_s / 2)))
    B = 2 if rate * target_s >= 2 * T else 1
    dt = min(run(B, T) for _ in range(2))
    return {"value": B * T / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle (torch CPU fp32 restatement of the reference modules), %d x %d samples, best of 2, "
                      "%.2f s" % (B, T, dt)}


def train_step_measure(w, rank, world, dist, max_over_ranks, barrier, steps=5, warmup=2):
    """The metric's "fwd+bwd" half: BASELINE configs[2], the WaveNet-CTC train step (legacy_code/train.py:24-61) on
    the same WaveNet plus the ecoli classifier, batch-sharded (weak scaling), bf16 tensor-core kernels with fp32
    master weights, Adam (wavenet_speech_b200.optim.Adam: torch.optim.Adam's update in one launch), gradient all-reduce when N > 1.  Inputs resident on the device."""
    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import train as TRN
    from wavenet_speech_b200.utils import signal_gen as S
    torch.manual_seed(0)
    C = w["C"]
    wn = W.WaveNet(w["in_dim"], w["entry_k"], [(C, C, 2, d) for d in w["dil"]], C, softmax=False).cuda()
    cn = W.WaveNetClassifier(C, 5, [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3], C, pool_kernel_size=3,
                             softmax=False).cuda()
    opt = W.optim.Adam(list(wn.parameters()) + list(cn.parameters()), lr=1e-5)    # torch.optim.Adam's rule, one launch
    B, T = w["batch"], w["T"]
    nb = min(B, 8)
    lev, labels = S.quantized_batch(nb, T, num_levels=w["in_dim"], seed=77 + rank, with_labels=True)
    rep = (B + nb - 1) // nb
    sig = torch.from_numpy(S.one_hot(lev, num_levels=w["in_dim"])).repeat(rep, 1, 1)[:B].cuda().bfloat16()
    labels = (labels * rep)[:B]
    nlab = T // 3 // 2                               # labels per read that an alignment over T/3 frames can carry
    lengths = torch.tensor([min(len(l), nlab) for l in labels], dtype=torch.int32)
    seq = (torch.cat([torch.from_numpy(l[:nlab]) for l in labels]) - 1).int().cuda()       # 0-based, as the reference

    def step():
        return TRN.train_step(wn, cn, sig, seq, lengths, opt, world=world)

    for _ in range(warmup):
        losses = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()          # (the collector stays ON here: 30 GB of activations per step must not wait for a cycle collection)
    e0.record()
    for _ in range(steps):
        losses = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    fl = 3 * (flops_per_timestep(w) + 16910848 / 3.0)
    return {"metric": "raw-signal samples/sec (WaveNet-CTC train step: fwd + bwd + Adam)", "value": world * B * T / (ms * 1e-3),
            "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "batch_per_gpu": B, "T": T,
            "classifier": "WaveNetClassifier 1+15 blocks, pool 3, 5 labels", "labels_per_read": int(lengths.max()),
            "tflops_as_written_3x_fwd": world * B * T / (ms * 1e-3) * fl / 1e12, "joint_loss": float(losses[2]),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}


def stock_torch_gpu(w, rank):
    """What a user of the reference runs on this GPU today: the reference's module semantics (modules/wavenet.py:88-111,
    block.py:54-82 -- nn.Conv1d padded + sliced, permute/contiguous + nn.Linear, F.tanh / F.sigmoid, per-layer add)
    issued as stock torch ops on `cuda` (cuDNN / cuBLAS / ATen), in fp32 (the reference's dtype) and in bf16.  The
    functional restatement under oracle/ issues exactly those ops; nothing of libwnb200 is on this path.  Forward on the
    full workload; fwd+bwd (cross-entropy + CTC through torch autograd, torch's own ctc_loss) on a bounded batch."""
    from oracle import wavenet_oracle as O
    import torch.nn.functional as F
    torch.manual_seed(0)
    net, layers = build_model(w)
    B, T, D = w["batch"], w["T"], w["in_dim"]
    x = make_input(w, rank)
    out = {}

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        sd = {k: v.detach().cuda().to(dt) for k, v in net.state_dict().items()}
        xd = x.cuda().to(dt)
        try:
            with torch.no_grad():
                ms = timed(lambda: O.wavenet_forward(sd, xd, layers, softmax=True), 3)
            out["fwd_" + name] = {"samples_per_s": B * T / (ms * 1e-3), "ms_per_step": ms, "batch": B, "T": T}
        except RuntimeError as e:                    # out of memory on a smaller card: say so, do not fail the bench
            out["fwd_" + name] = {"error": str(e)[:120]}
        del sd, xd
        torch.cuda.empty_cache()
    # train step (legacy_code/train.py:24-61 semantics) with stock autograd, bf16, bounded batch
    try:
        from wavenet_speech_b200.utils import signal_gen as S
        Bt = min(B, 4)
        C = w["C"]
        import wavenet_speech_b200 as W
        cls_layers = [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3]
        cn = W.WaveNetClassifier(C, 5, cls_layers, C, pool_kernel_size=3, softmax=False)
        sd_w = {k: v.detach().cuda().bfloat16().requires_grad_(True) for k, v in net.state_dict().items()}
        sd_c = {k: v.detach().cuda().bfloat16().requires_grad_(True) for k, v in cn.state_dict().items()}
        lev, labels = S.quantized_batch(Bt, T, num_levels=D, seed=77, with_labels=True)
        sig = torch.from_numpy(S.one_hot(lev, num_levels=D)).cuda().bfloat16()
        nlab = T // 3 // 2
        lengths = torch.tensor([min(len(l), nlab) for l in labels], dtype=torch.int64)
        seq = torch.cat([torch.from_numpy(l[:nlab]) for l in labels]).long().cuda()
        params = list(sd_w.values()) + list(sd_c.values())
        opt = torch.optim.Adam(params, lr=1e-5, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            pred = O.wavenet_forward(sd_w, sig[:, :, 0:-1], layers, softmax=False)
            trans = O.classifier_forward(sd_c, pred, cls_layers, pool_kernel_size=3, softmax=False)
            tgt = sig[:, :, 1:].argmax(1)
            xe = F.cross_entropy(pred.float(), tgt, reduction="sum") / Bt          # sum over t of batch means
            lp = F.log_softmax(trans.float().permute(2, 0, 1), 2)
            ctc = F.ctc_loss(lp, seq, torch.full((Bt,), lp.shape[0], dtype=torch.int64), lengths, blank=0,
                             reduction="sum")
            (xe / T + ctc / lp.shape[0]).backward()
            opt.step()

        ms = timed(step, 2)
        out["train_step_bf16"] = {"samples_per_s": Bt * T / (ms * 1e-3), "ms_per_step": ms, "batch": Bt, "T": T,
                                  "note": "stock autograd + torch ctc_loss + fused Adam; bounded batch (autograd keeps "
                                          "every intermediate of every layer)"}
    except RuntimeError as e:
        out["train_step_bf16"] = {"error": str(e)[:120]}
    torch.cuda.empty_cache()
    return out


def long_read_measure(rank, world, dist, max_over_ranks, barrier, T=1000000, steps=5):
    """BASELINE configs[4]: one 1M-sample read through the ecoli RawCTCNet, time-sharded over the ranks with the
    receptive-field halo (51 / 45 samples) exchanged over NVLink (NCCL point-to-point); every rank checks its span
    bit for bit against the single-pass output."""
    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import sharding as S
    from wavenet_speech_b200.utils import signal_gen as SG
    torch.manual_seed(0)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in [1, 2, 4, 8, 16] * 3], 256, softmax=False)
    net = net.cuda().bfloat16().eval()
    xfull = torch.from_numpy(SG.raw_batch(1, T, seed=11)).bfloat16()
    hl, hr = S.raw_ctcnet_halo(net)
    plan = S.time_shard_plan(T, rank, world, hl, hr)
    mine = xfull[:, :, plan["start"]:plan["end"]].contiguous().cuda()
    extra = net.feature_kwidth - 1
    out = {}

    def run():
        x_ext = S.exchange_halo(mine, plan, rank, world) if world > 1 else mine
        out["y"] = S.time_sharded_forward(net, x_ext, plan, T, out_extra=extra)

    def halo_only():
        if world > 1:
            S.exchange_halo(mine, plan, rank, world)

    with torch.no_grad():
        for _ in range(2):
            run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        e0.record()
        for _ in range(steps):
            halo_only()
        e1.record()
        barrier()
        ms_halo = max_over_ranks(e0.elapsed_time(e1)) / steps
        full = net(xfull.cuda())
        stop = plan["end"] + (extra if plan["end"] == T else 0)
        ok = torch.tensor([int(torch.equal(out["y"], full[:, :, plan["start"]:stop]))], device="cuda")
        if dist is not None:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"workload": "rawctcnet_ecoli_fk3_1x%d_time_sharded" % T, "samples_per_s": T / (ms * 1e-3),
            "ms_per_read": ms, "halo": [hl, hr], "halo_bytes_per_boundary": 2 * (hl + hr),
            "exchange": "nccl p2p (batched isend/irecv over NVLink)" if world > 1 else "none (1 rank)",
            "halo_exchange_ms": ms_halo, "sharded_equals_full_bitwise": bool(int(ok)), "scaling": "strong"}


def rawctc_sweep(rank, world, dist, max_over_ranks, barrier, steps=3):
    """BASELINE configs[3]: ecoli RawCTCNet (configs/ecoli_testrun.json: 256 channels, input block + (1,2,4,8,16) x 3,
    fk = 3) forward throughput for global batches 64 ... 1024 x 4000 samples, the batch sharded over the ranks (no
    collective: reads are independent).  bf16 in/out, precise format."""
    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import sharding as S
    from wavenet_speech_b200.utils import signal_gen as SG
    torch.manual_seed(0)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in [1, 2, 4, 8, 16] * 3], 256, softmax=False)
    net = net.cuda().bfloat16().eval()
    base = torch.from_numpy(SG.raw_batch(64, 4000, seed=7 + rank)).bfloat16().cuda()
    out = {}
    with torch.no_grad():
        for Bg in (64, 256, 1024):
            s0, s1 = S.shard_range(Bg, rank, world)
            n = s1 - s0
            x = base.repeat((n + 63) // 64, 1, 1)[:n].contiguous() if n > 0 else base[:0]
            for _ in range(2):
                net(x)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                net(x)
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / steps
            out["B%d" % Bg] = {"samples_per_s": Bg * 4000 / (ms * 1e-3), "ms_per_step": ms}
    out["workload"] = "rawctcnet_ecoli_fk3 forward, global batch x 4000 samples, batch-sharded x%d" % world
    return out


def run_reference(args, w):
    """--impl reference: the reference is Python/torch and cannot travel to the GPU box (and is not
    pip-installable: it has no setup.py), so this arm times the oracle port of its CPU path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import wavenet_oracle as O
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    net, layers = build_model(w)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    D = w["in_dim"]
    # per step: 2 full-length reads of the workload (the same sample `cpu_baseline` uses).  Shorter reads flatter the CPU:
    # a 4096-sample read keeps the layer tensors (256 ch x T x 4 B) in cache and runs ~2x faster per sample than the
    # 16384-sample reads the workload is made of.
    B, T = min(2, w["batch"]), w["T"]
    lev = torch.randint(0, D, (B, T))
    x = torch.zeros(B, D, T).scatter_(1, lev.unsqueeze(1), 1.0)
    with torch.no_grad():
        for _ in range(args.warmup):
            O.wavenet_forward(sd, x, layers, softmax=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.wavenet_forward(sd, x, layers, softmax=True)
        dt = time.perf_counter() - t0
    val = B * T * args.steps / dt
    sample = "oracle port, fp32, %d x %d samples per step (bounded sample of the %d x %d workload)" % (
        B, T, w["batch"], w["T"])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the fwd+bwd (config 3 train step) measurement")
    ap.add_argument("--e2e-chunks", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true",
                    help="device-resident timing: issue the forward's launches from Python instead of one graph replay per step")
    ap.add_argument("--no-e2e-graph", action="store_true",
                    help="e2e: issue every chunk's ~25 launches from Python instead of replaying one CUDA graph per chunk")
    ap.add_argument("--precision", default="precise", choices=["precise", "fast"],
                    help="tensor-core activation format: precise = fp16 operands + fp16 (hi, lo) residual stream + exact "
                         "gate (meets 2e-2 at 20 blocks); fast = bf16 stream + tanh.approx (round-1 format)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock-PyTorch-on-GPU comparator")
    ap.add_argument("--no-longread", action="store_true", help="skip the time-sharded 1M-sample read (config 5)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    if args.T:
        w["T"] = args.T
    if args.impl == "reference":
        run_reference(args, w)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    if not args.no_numa_bind:          # before any pinned allocation: first touch places the staging buffers
        from wavenet_speech_b200.utils import numa as _numa
        numa = _numa.bind_to_gpu_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import _lib
    W.tc_precision(args.precision)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    net, layers = build_model(w)
    net = net.cuda().to(dtype).eval()
    x_host = make_input(w, rank).to(dtype)
    x_host = x_host.pin_memory()
    x_dev = x_host.cuda(non_blocking=True)
    samples = w["batch"] * w["T"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    def graphed(tag):
        """The forward as ONE CUDA-graph replay per step (pipeline.GraphedForward, captured on x_dev itself): a 20-step
        region that starts from an empty launch queue otherwise depends on the Python thread issuing ~25 launches per
        step without ever losing 50 ms (torchrun, 2 ranks: 74 M samples/s read where 93 M is the rate).  Same kernels,
        same order; --no-graph issues them from Python."""
        if args.no_graph:
            return None
        try:
            from wavenet_speech_b200.pipeline import GraphedForward
            return GraphedForward(net, x_dev, static_input=True)
        except Exception as e:                                   # (a path that cannot be captured: time it eagerly)
            sys.stderr.write("bench: %s forward not captured (%s); eager launches\n" % (tag, e))
            return None

    with torch.no_grad():
        for _ in range(args.warmup):
            y = net(x_dev)
        torch.cuda.synchronize()
        l0 = _lib.launch_count
        y = net(x_dev)
        launches_per_forward = _lib.launch_count - l0            # our kernels in one forward (eager count)
        gf = graphed("timed")
        if gf is not None:
            for _ in range(2):
                gf.graph.replay()
            y = gf.y
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with quiet_gc():
            w0 = time.time()
            e0.record()
            for _ in range(args.steps):
                if gf is not None:
                    gf.graph.replay()
                else:
                    y = net(x_dev)
            e1.record()
            barrier()
        sampler.mark(w0, time.time())
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = launches_per_forward * args.steps             # executed inside the timed region (in the graph or not)
        clocks = sampler.stop()
        launch_mode = "cuda graph replay (pipeline.GraphedForward), 1 launch per step" if gf is not None else "eager"

        # ---- the other activation format on the same input, same steps (reported under config) -----------
        other = "fast" if args.precision == "precise" else "precise"
        other_fmt = None
        if args.dtype == "bf16":
            with W.tc_precision(other):
                for _ in range(2):
                    net(x_dev)
                gfo = graphed("other-format")
                if gfo is not None:
                    gfo.graph.replay()
                barrier()
                with quiet_gc():
                    e0.record()
                    for _ in range(args.steps):
                        if gfo is not None:
                            gfo.graph.replay()
                        else:
                            net(x_dev)
                    e1.record()
                    barrier()
                del gfo
            ms_o = max_over_ranks(e0.elapsed_time(e1))
            other_fmt = {"precision": other, "value": samples * world * args.steps / (ms_o * 1e-3), "unit": UNIT,
                         "ms_per_step": ms_o / args.steps}

        # ---- end-to-end: pinned host input -> H2D -> forward -> D2H of the output ------------------------
        e2e = None
        if not args.no_e2e:
            from wavenet_speech_b200.pipeline import HostPipeline
            y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            del y
            pipe = HostPipeline(net, chunks=args.e2e_chunks, graph=not args.no_e2e_graph)   # public API: pinned host in -> pinned host out
            y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
            # warm-up = the timed loop's own pattern (back-to-back submits): the caching allocator must have seen the
            # overlap of step k's copy-out with step k+1's kernels, or it grows its pool with (synchronising) cudaMalloc
            # calls INSIDE the timed region -- measured: e2e 21-28 M samples/s on the runs where that happened, 43-45 M
            # otherwise, same binary, same box
            for k in range(2):
                pipe(x_host, y_hosts[k])
            for k in range(max(3, args.warmup) + 2):
                pipe.submit(x_host, y_hosts[k & 1])
            pipe.wait()
            barrier()
            n_malloc0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
            with quiet_gc():
                t0 = time.perf_counter()
                e0.record()
                for k in range(args.steps):                      # every step: H2D of its input, kernels, D2H of its
                    pipe.submit(x_host, y_hosts[k & 1])          # output; steps are streamed back to back
                pipe.wait()                                      # ... and the last D2H copy has landed
                e1.record()
                barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
            e2e = {"value": samples * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
                   "d2h_bytes_per_step": y_host.numel() * y_host.element_size(),
                   "ms_per_step": ms_e2e / args.steps,
                   "cuda_mallocs_in_timed_region": torch.cuda.memory_stats().get("num_device_alloc", 0) - n_malloc0,
                   "api": "wavenet_speech_b200.pipeline.HostPipeline(model, chunks=%d, graph=%s).submit(x_pinned, y_pinned) per "
                          "step, .wait() once at the end" % (args.e2e_chunks, not args.no_e2e_graph)}

        # ---- what the HOST can deliver: every rank moves the step's bytes in and out (same pinned buffers, same two
        # copy streams) with no kernels at all.  e2e cannot be faster than this; at N = 8 it is the limiter.
        e2e_ceiling = None
        if not args.no_e2e:
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            xd2 = torch.empty_like(x_dev)
            yd2 = torch.empty(y_host.shape, dtype=y_host.dtype, device="cuda")
            for _ in range(2):
                xd2.copy_(x_host, non_blocking=True)
                y_hosts[0].copy_(yd2, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for k in range(args.steps):
                with torch.cuda.stream(s_in):
                    xd2.copy_(x_host, non_blocking=True)
                with torch.cuda.stream(s_out):
                    y_hosts[k & 1].copy_(yd2, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            barrier()
            ms_c = max_over_ranks((time.perf_counter() - t0) * 1e3)
            nbytes = (x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size())
            e2e_ceiling = {"value": samples * world * args.steps / (ms_c * 1e-3), "unit": UNIT,
                           "ms_per_step": ms_c / args.steps,
                           "host_gb_per_s_all_ranks": world * nbytes * args.steps / (ms_c * 1e-3) / 1e9,
                           "what": "H2D of the input + D2H of the output per step on two copy streams, no kernels: "
                                   "the copy ceiling of this host at this N"}
            del xd2, yd2

        # ---- the same end to end, with the host handing over the quantised LEVELS (B, T) uint8 instead of their
        # one-hot encoding: what the reference's loaders hold before fns.py:6-15; the one-hot never exists on this path
        # (the entry conv is a gather).  Extra information, not the contract's `e2e` (which keeps the reference call).
        e2e_levels = None
        if not args.no_e2e and args.dtype == "bf16" and w["in_dim"] == w["C"] and w["C"] in (128, 256):
            from wavenet_speech_b200.pipeline import HostPipeline
            lev_host = x_host.float().argmax(1).to(torch.uint8).pin_memory()
            pipe2 = HostPipeline(net, chunks=args.e2e_chunks, fn=net.forward_levels, graph=not args.no_e2e_graph)
            for k in range(2):
                pipe2(lev_host, y_hosts[k])
            for k in range(max(3, args.warmup) + 2):
                pipe2.submit(lev_host, y_hosts[k & 1])
            pipe2.wait()
            barrier()
            n_malloc0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
            with quiet_gc():
                t0 = time.perf_counter()
                e0.record()
                for k in range(args.steps):
                    pipe2.submit(lev_host, y_hosts[k & 1])
                pipe2.wait()
                e1.record()
                barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_l = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
            e2e_levels = {"value": samples * world * args.steps / (ms_l * 1e-3), "unit": UNIT,
                          "cuda_mallocs_in_timed_region": torch.cuda.memory_stats().get("num_device_alloc", 0) - n_malloc0,
                          "h2d_bytes_per_step": lev_host.numel(), "d2h_bytes_per_step": y_host.numel() * y_host.element_size(),
                          "ms_per_step": ms_l / args.steps,
                          "api": "HostPipeline(model, chunks=%d, fn=model.forward_levels).submit(levels_u8_pinned, y_pinned)"
                                 % args.e2e_chunks}

        # ---- roofline of the dominant kernel: per-launch CUDA-event timing on the launching stream ------
        # (a second pass of the same K steps, so that the per-launch average sees the same sustained clocks as `value`:
        # the first steps after an idle period run ~10 % faster, before the power limiter pulls the SM clock down)
        _lib.kernel_timing(True)
        for _ in range(args.steps):
            net(x_dev)
        torch.cuda.synchronize()
        log = _lib.kernel_timing(False)
    n_pass = args.steps
    per = {}
    for name, a, b in log:
        per.setdefault(name, []).append(a.elapsed_time(b))
    peaks = load_peaks()
    C = w["C"]
    precise = args.precision == "precise" and args.dtype == "bf16" and C in (128, 256)
    tagged = [k for k in per if k.endswith(":resblock")]
    dom = tagged[0] if tagged else "wnb200_taps_fwd"
    tot = sum(sum(v) for v in per.values())
    from wavenet_speech_b200 import fastpath as _FP
    deferred = bool(tagged) and _FP.DEFER_SKIP and any(k.endswith(":skipsum") for k in per)
    L = len(w["dil"])
    if tagged and deferred:
        # deferred skip: a block launch computes the two dilated convs, conv1x1_residual and residual_proj (12 C^2 FLOP per
        # frame as written, block.py:66-79; the last layer has no residual output: 8 C^2); conv1x1_skip + the bottleneck
        # (4 C^2 per layer as written, block.py:74 + wavenet.py:100) are the stack-wide skip contraction, reported apart
        flops_launch = ((L - 1) * 12 + 8) / float(L) * C * C * samples
        avg_ms = float(np.mean(per[dom]))
    elif tagged:
        flops_launch = 16 * C * C * samples                        # block + bottleneck as written
        avg_ms = float(np.mean(per[dom]))
    else:                                                           # generic path: all contraction launches together
        flops_launch = flops_per_timestep(w) * samples
        avg_ms = float(sum(per[dom])) / float(n_pass)
    achieved = flops_launch / (avg_ms * 1e-3) / 1e12
    # DRAM bytes per launch of the fused block kernel: read from the ncu summary of THIS kernel / format / workload
    # that scripts/ncu_traffic.py wrote under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum of one
    # `ncu --set full` launch); null when no capture of this format exists -- never a constant in this file
    traffic, traffic_src = None, None
    if tagged and args.workload == DEFAULT_WORKLOAD and not args.batch and not args.T:
        tp = os.path.join(ROOT, "profiles", "r2_resblock_traffic.json")
        if os.path.exists(tp):
            ent = json.load(open(tp)).get(args.precision)
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent.get("source")
    # executed MACs per frame: bf16 format 7 C^2 (skip -> bottleneck folded); precise 8 C^2 (+ the stream's lo half
    # through the projection); as written in the reference: 8 C^2 (16 C^2 FLOP)
    if deferred:       # executed: + the lo half through the projection in the precise format (14 of 12 as written)
        exec_ratio = (((L - 1) * 14 + 8) / float((L - 1) * 12 + 8)) if precise else 1.0
    else:
        exec_ratio = (16.0 if precise else 14.0) / 16.0
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic,
                "traffic_source": traffic_src,
                "achieved_executed": achieved * exec_ratio if tagged else achieved,
                "peak_source": peaks["source"] + " (bf16_tflops_sustained)",
                "share_of_step": float(sum(per[dom])) / tot if tot > 0 else None,
                "avg_launch_ms": avg_ms,
                "flops_per_launch_as_written": flops_launch}
    if deferred:
        sk = [k for k in per if k.endswith(":skipsum")][0]
        sk_ms = float(np.mean(per[sk]))
        roofline["skip_contraction"] = {
            "kernel": sk, "avg_launch_ms": sk_ms, "launches_per_step": len(per[sk]) // n_pass,
            "tflops_as_written": 4 * C * C * L * samples / (sk_ms * 1e-3) / 1e12,
            "tflops_executed": 2 * C * C * L * samples / (sk_ms * 1e-3) / 1e12,
            "what": "sum_l (Wbn_l Wskip_l) gate_l over K = layers x channels, accumulated in TMEM (wavenet.py:97-100)"}
        # the residual stack as a whole (SURVEY 8d's 16 C^2 FLOP per frame and layer as written = the block launches plus
        # the skip contraction), over the summed launch times of one pass
        stack_ms = (float(sum(per[dom])) + float(sum(per[sk]))) / float(n_pass)
        stack_tf = 16 * C * C * L * samples / (stack_ms * 1e-3) / 1e12
        roofline["stack"] = {"ms_per_step": stack_ms, "tflops_as_written": stack_tf,
                             "frac": stack_tf / peaks["bf16_tflops_sustained"],
                             "what": "all block launches + the skip contraction of one forward against 16 C^2 FLOP per frame "
                                     "and layer as written (block.py:54-82 + wavenet.py:100)"}

    input_mb = x_dev.numel() * x_dev.element_size() / 1e6
    fwd_bwd = None
    del net, x_dev
    torch.cuda.empty_cache()
    if not args.no_train and args.workload == DEFAULT_WORKLOAD and args.dtype == "bf16":
        fwd_bwd = train_step_measure(w, rank, world, dist, max_over_ranks, barrier)
        torch.cuda.empty_cache()
    long_read = None
    if not args.no_longread and args.workload == DEFAULT_WORKLOAD and args.dtype == "bf16":
        long_read = long_read_measure(rank, world, dist, max_over_ranks, barrier)
        torch.cuda.empty_cache()
    sweep = None
    if not args.no_longread and args.workload == DEFAULT_WORKLOAD and args.dtype == "bf16":
        sweep = rawctc_sweep(rank, world, dist, max_over_ranks, barrier)
        torch.cuda.empty_cache()
    stock = None
    if not args.no_stock and world == 1 and args.workload == DEFAULT_WORKLOAD:
        stock = stock_torch_gpu(w, rank)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None if args.no_cpu_baseline else cpu_baseline(w)
    value = samples * world * args.steps / (ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": w["batch"], "T": w["T"], "channels": C,
                   "layers": len(w["dil"]), "softmax": True, "sharding": "batch x%d" % world,
                   "l2": "input %.0f MB and every inter-layer tensor exceed the 126 MB L2" % input_mb,
                   "flop_per_sample": flops_per_timestep(w),
                   "launch": launch_mode,
                   "precision": ("precise: bf16 in/out, fp16 tensor-core operands, residual stream as an fp16 (hi, lo) pair, "
                                 "fp32 accumulation (<= 2e-2 at 20 blocks, tests/test_gpu_precise.py)") if precise else
                                "fast: bf16 operands and residual stream, tanh.approx gate",
                   # the rest of BASELINE.json's metric, measured in this same run at this N (the driver keeps `config`):
                   "other_format": other_fmt,
                   "fwd_bwd": fwd_bwd,                 # configs[2]: WaveNet-CTC train step, batch-sharded + grad all-reduce
                   "time_sharded": long_read,          # configs[4]: 1M-sample read, time-sharded + halo exchange
                   "rawctcnet_batch_sweep": sweep,     # configs[3]: ecoli RawCTCNet forward, batch 64 .. 1024 x 4000
                   "e2e_levels": e2e_levels,           # host hands over uint8 levels instead of the one-hot tensor
                   "e2e_copy_ceiling": e2e_ceiling,    # pure H2D + D2H of the same bytes at this N: the host's limit
                   "e2e_frac_of_copy_ceiling": (e2e["value"] / e2e_ceiling["value"]) if (e2e and e2e_ceiling) else None,
                   "numa": numa,
                   "stock_torch_gpu": stock},          # the reference's module semantics as stock torch ops on this GPU
        "clocks": clocks, "e2e": e2e, "e2e_levels": e2e_levels, "gpu_launches": launches, "roofline": roofline,
        "cpu_baseline": cpu,
        "fwd_bwd": fwd_bwd,
        "tflops": value * flops_per_timestep(w) / 1e12,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
