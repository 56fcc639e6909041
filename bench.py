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
    master weights, fused Adam, gradient all-reduce when N > 1.  Inputs resident on the device."""
    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import train as TRN
    from wavenet_speech_b200.utils import signal_gen as S
    torch.manual_seed(0)
    C = w["C"]
    wn = W.WaveNet(w["in_dim"], w["entry_k"], [(C, C, 2, d) for d in w["dil"]], C, softmax=False).cuda()
    cn = W.WaveNetClassifier(C, 5, [(C, C, 2, d) for d in [1, 2, 4, 8, 16] * 3], C, pool_kernel_size=3,
                             softmax=False).cuda()
    opt = torch.optim.Adam(list(wn.parameters()) + list(cn.parameters()), lr=1e-5, fused=True)
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import wavenet_speech_b200 as W
    from wavenet_speech_b200 import _lib
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
    with torch.no_grad():
        for _ in range(args.warmup):
            y = net(x_dev)
        barrier()
        l0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(args.steps):
            y = net(x_dev)
        e1.record()
        barrier()
        sampler.mark(w0, time.time())
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = _lib.launch_count - l0
        clocks = sampler.stop()

        # ---- end-to-end: pinned host input -> H2D -> forward -> D2H of the output ------------------------
        e2e = None
        if not args.no_e2e:
            from wavenet_speech_b200.pipeline import HostPipeline
            y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            del y
            pipe = HostPipeline(net, chunks=args.e2e_chunks)     # public API: pinned host in -> pinned host out
            y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
            for k in range(2):
                pipe(x_host, y_hosts[k])
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for k in range(args.steps):                          # every step: H2D of its input, kernels, D2H of its
                pipe.submit(x_host, y_hosts[k & 1])              # output; steps are streamed back to back
            pipe.wait()                                          # ... and the last D2H copy has landed
            e1.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
            e2e = {"value": samples * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
                   "d2h_bytes_per_step": y_host.numel() * y_host.element_size(),
                   "ms_per_step": ms_e2e / args.steps,
                   "api": "wavenet_speech_b200.pipeline.HostPipeline(model, chunks=%d).submit(x_pinned, y_pinned) per step, "
                          ".wait() once at the end" % args.e2e_chunks}

        # ---- the same end to end, with the host handing over the quantised LEVELS (B, T) uint8 instead of their
        # one-hot encoding: what the reference's loaders hold before fns.py:6-15; the one-hot never exists on this path
        # (the entry conv is a gather).  Extra information, not the contract's `e2e` (which keeps the reference call).
        e2e_levels = None
        if not args.no_e2e and args.dtype == "bf16" and w["in_dim"] == w["C"] and w["C"] in (128, 256):
            from wavenet_speech_b200.pipeline import HostPipeline
            lev_host = x_host.float().argmax(1).to(torch.uint8).pin_memory()
            pipe2 = HostPipeline(net, chunks=args.e2e_chunks, fn=net.forward_levels)
            for k in range(2):
                pipe2(lev_host, y_hosts[k])
            barrier()
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
    tagged = [k for k in per if k.endswith(":resblock")]
    dom = tagged[0] if tagged else "wnb200_taps_fwd"
    tot = sum(sum(v) for v in per.values())
    if tagged:
        flops_launch = 16 * C * C * samples                        # block + bottleneck as written
        avg_ms = float(np.mean(per[dom]))
    else:                                                           # generic path: all contraction launches together
        flops_launch = flops_per_timestep(w) * samples
        avg_ms = float(sum(per[dom])) / float(n_pass)
    achieved = flops_launch / (avg_ms * 1e-3) / 1e12
    # DRAM bytes per launch of the fused block kernel from the committed `ncu --set full` capture of this same
    # workload (profiles/r1_ncu_full_v8_resblock2.csv): 806.9 MB read + 749.9 MB written
    traffic = 1556.8e6 if (tagged and args.workload == DEFAULT_WORKLOAD and not args.batch and not args.T) else None
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic,
                "achieved_executed": achieved * 14.0 / 16.0 if tagged else achieved,
                "peak_source": peaks["source"] + " (bf16_tflops_sustained)",
                "share_of_step": float(sum(per[dom])) / tot if tot > 0 else None,
                "avg_launch_ms": avg_ms}

    input_mb = x_dev.numel() * x_dev.element_size() / 1e6
    fwd_bwd = None
    if not args.no_train and args.workload == DEFAULT_WORKLOAD and args.dtype == "bf16":
        del net, x_dev
        torch.cuda.empty_cache()
        fwd_bwd = train_step_measure(w, rank, world, dist, max_over_ranks, barrier)
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
                   "flop_per_sample": flops_per_timestep(w)},
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
