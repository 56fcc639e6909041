"""Multi-GPU measurements of the two partitionings (BASELINE configs 3, 4, 5); one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_multi.py [train] [longread] [rawctc] [B=32] [T=16384]

Every time is the MAX over ranks of a CUDA-event interval bracketed by barrier + synchronize; rank 0 prints one JSON
line per measurement.  With N = 1 the same code runs without collectives (the single-GPU reference point).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import wavenet_speech_b200 as W
from wavenet_speech_b200 import sharding as S
from wavenet_speech_b200.utils import signal_gen as SG

which = set(a for a in sys.argv[1:] if "=" not in a) or {"train", "longread", "rawctc"}
opts = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ECOLI = [1, 2, 4, 8, 16] * 3


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def emit(d):
    if rank == 0:
        d["n_gpus"] = world
        print(json.dumps(d), flush=True)


if "train" in which:
    # config 3: WaveNet NLL + classifier CTC train step, batch-sharded (B reads per GPU: weak scaling), one bucketed
    # NCCL gradient all-reduce per step; cross-entropy pre-scaled by 1/world (batch mean), CTC not (batch sum).
    torch.manual_seed(0)
    dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2
    wn = W.WaveNet(256, 2, [(256, 256, 2, d) for d in dil], 256, softmax=False).cuda()
    cn = W.WaveNetClassifier(256, 5, [(256, 256, 2, d) for d in ECOLI], 256, pool_kernel_size=3, softmax=False).cuda()
    params = list(wn.parameters()) + list(cn.parameters())
    opt = torch.optim.Adam(params, lr=1e-5, fused=True)
    B, T = int(opts.get("B", 32)), int(opts.get("T", 16384))
    lev, labels = SG.quantized_batch(min(B, 8), T, seed=5 + rank, with_labels=True)
    rep = (B + 7) // 8
    sig = torch.from_numpy(SG.one_hot(lev)).repeat(rep, 1, 1)[:B].cuda().bfloat16()
    labels = (labels * rep)[:B]
    nlab = T // 3 // 2
    lengths = torch.tensor([min(len(l), nlab) for l in labels], dtype=torch.int32)
    seq = torch.cat([torch.from_numpy(l[:nlab]) for l in labels]).int().cuda()
    stats = {}

    def step():
        opt.zero_grad(set_to_none=True)
        pred = wn(sig[:, :, 0:-1])
        trans = cn(pred)
        dense = W.ops.argmax_channels(sig[:, :, 1:].contiguous())
        xe = W.functional.cross_entropy_sum(pred, dense) / B
        ctc = W.functional.ctc_loss_sum(trans, seq, lengths, layout="bct")
        loss = S.joint_loss_for_backward(xe, ctc, T, trans.shape[2], world)
        loss.backward()
        stats["buckets"] = S.allreduce_gradients(params)
        opt.step()
        return loss

    l0 = float(step())
    ms = timed(step, int(opts.get("steps", 5)))
    # replicas must stay identical: compare a checksum of the parameters across ranks
    chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
    same = True
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(float(c) == float(allc[0]) for c in allc)
    nbytes = sum(p.numel() * 4 for p in params if p.grad is not None)
    emit({"config": "wavenet_ctc_train_step_bf16_tc_batch_sharded", "batch_per_gpu": B, "T": T, "ms_per_step": ms,
          "samples_per_s": world * B * T / (ms * 1e-3), "scaling": "weak", "allreduce_bytes": nbytes,
          "allreduce_buckets": stats.get("buckets", 0), "replicas_identical": same, "loss_first": l0,
          "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
    del wn, cn, opt, params, sig
    torch.cuda.empty_cache()

if "rawctc" in which:
    # config 4: ecoli RawCTCNet, bf16 inference, global batch sharded over the ranks (strong scaling, no collective)
    torch.manual_seed(0)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in ECOLI], 256, softmax=False).cuda().bfloat16().eval()
    for Bg in (256, 1024):
        s0, s1 = S.shard_range(Bg, rank, world)
        x = torch.from_numpy(SG.raw_batch(64, 4000, seed=7)).repeat((s1 - s0 + 63) // 64, 1, 1)[:s1 - s0].cuda().bfloat16()
        with torch.no_grad():
            ms = timed(lambda: net(x), 5)
        emit({"config": "rawctcnet_ecoli_fk3_bf16_fwd_batch_sharded", "global_batch": Bg, "T": 4000, "ms_per_step": ms,
              "samples_per_s": Bg * 4000 / (ms * 1e-3), "scaling": "strong"})

if "longread" in which:
    # config 5: one 1M-sample read, time-sharded; each rank holds ITS span only, gets the receptive-field halo from
    # its neighbours over NCCL point-to-point (NVLink), recomputes the fringe and keeps its own output span.
    torch.manual_seed(0)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in ECOLI], 256, softmax=False).cuda().bfloat16().eval()
    T = int(opts.get("TL", 1000000))
    xfull = torch.from_numpy(SG.raw_batch(1, T, seed=11)).bfloat16()          # same on every rank (same seed)
    hl, hr = S.raw_ctcnet_halo(net)
    plan = S.time_shard_plan(T, rank, world, hl, hr)
    mine = xfull[:, :, plan["start"]:plan["end"]].contiguous().cuda()
    extra = net.feature_kwidth - 1
    out = {}

    def run():
        x_ext = S.exchange_halo(mine, plan, rank, world) if world > 1 else mine
        out["y"] = S.time_sharded_forward(net, x_ext, plan, T, out_extra=extra)

    peer = None
    if world > 1 and opts.get("halo", "both") != "nccl":      # NVLink peer-memory exchange (no NCCL on the data path)
        try:
            peer = S.PeerHaloExchange(1, 1, hl, hr, mine.dtype, rank, world)
        except Exception as e:                                 # symmetric memory unavailable on this box
            emit({"config": "peer_halo_exchange_unavailable", "error": str(e)[:200]})

    def run_peer():
        out["y"] = S.time_sharded_forward(net, peer.exchange(mine, plan), plan, T, out_extra=extra)

    with torch.no_grad():
        if peer is not None:
            ms_peer = timed(run_peer, 5)
            y_peer = out["y"].clone()
        ms = timed(run, 5)
        if peer is not None:
            same = torch.tensor([int(torch.equal(out["y"], y_peer))], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            emit({"config": "rawctcnet_ecoli_1M_read_time_sharded_peer_memory_halo", "T": T, "ms_per_read": ms_peer,
                  "samples_per_s": T / (ms_peer * 1e-3), "equals_nccl_exchange_bitwise": bool(int(same))})
        # exactness: every rank checks its span against the full read computed locally in one pass
        full = net(xfull.cuda())
        stop = plan["end"] + (extra if plan["end"] == T else 0)
        ok = torch.tensor([int(torch.equal(out["y"], full[:, :, plan["start"]:stop]))], device="cuda")
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    emit({"config": "rawctcnet_ecoli_1M_read_time_sharded", "T": T, "halo": [hl, hr], "ms_per_read": ms,
          "samples_per_s": T / (ms * 1e-3), "sharded_equals_full_bitwise": bool(int(ok)),
          "halo_bytes_per_boundary": 2 * (hl + hr), "scaling": "strong"})

if "peercausal" in which and world > 1:
    # one-directional halos (any causal network: halo_right = 0).  A rank only WRITES to its right neighbour
    # and never waits on it, so without the explicit ack a fast writer could rewrite a parity slot two steps later while the
    # reader still held step k.  Eight steps with a different read each, the reader artificially slowed on odd ranks; every
    # step must equal the NCCL exchange bit for bit.
    torch.manual_seed(0)
    net = W.RawCTCNet(256, 3, 5, [(256, 256, 2, d) for d in (1, 2, 4, 8)], 256, softmax=False, causal=True)
    net = net.cuda().bfloat16().eval()
    T = 200000
    hl, hr = S.raw_ctcnet_halo(net)
    assert hr == 0 and hl > 0
    plan = S.time_shard_plan(T, rank, world, hl, hr)
    peer = S.PeerHaloExchange(1, 1, hl, hr, torch.bfloat16, rank, world)
    ok = True
    with torch.no_grad():
        mines = [torch.from_numpy(SG.raw_batch(1, T, seed=100 + step)).bfloat16()[:, :, plan["start"]:plan["end"]]
                 .contiguous().cuda() for step in range(8)]
        got = []
        for step in range(8):                              # nothing but the protocol's own signals orders the ranks here
            if rank % 2 == 1:
                torch.cuda._sleep(20_000_000)              # a slow reader: ~10 ms late into every step
            got.append(peer.exchange(mines[step], plan).clone())
        for step in range(8):
            ok = ok and torch.equal(got[step], S.exchange_halo(mines[step], plan, rank, world))
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    emit({"config": "peer_halo_exchange_causal_one_directional_8_steps", "halo": [hl, hr],
          "equals_nccl_exchange_bitwise_every_step": bool(int(flag))})

if world > 1:
    dist.destroy_process_group()
