"""Host-side logic of the two multi-GPU partitionings, on CPU with the gloo backend (world size 2); the network
itself is stood in for by the oracle (the product has no CPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wavenet_speech_b200 as W
from wavenet_speech_b200 import sharding as S
from oracle import wavenet_oracle as O


def test_shard_range_and_halos():
    assert [S.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert S.shard_range(2, 3, 4) == (2, 2)
    wn = W.WaveNet(8, 2, [(8, 8, 2, d) for d in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2], 8)
    assert S.wavenet_halo(wn) == (2047, 0)
    rn = W.RawCTCNet(8, 3, 5, [(8, 8, 2, d) for d in [1, 2, 4, 8, 16] * 3], 8, softmax=False)
    assert S.raw_ctcnet_halo(rn) == (51, 45)
    p = S.time_shard_plan(1000, 1, 4, 51, 45)
    assert p == {"start": 250, "end": 500, "lo": 199, "hi": 545, "halo_left": 51, "halo_right": 45}
    p = S.time_shard_plan(1000, 0, 4, 51, 45, align=3)
    assert p["start"] == 0 and p["end"] == 252 and p["lo"] == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        # ---------------- time sharding: RawCTCNet over one long read, halo exchange, exact result ----------
        torch.manual_seed(7)
        layers = [(8, 8, 2, d) for d in (1, 2, 4, 8)] + [(8, 8, 3, 2)]
        net = W.RawCTCNet(8, 3, 5, layers, 8, softmax=False)
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        T = 301
        x = torch.randn(2, 1, T)
        full = O.raw_ctcnet_forward(sd, x, layers, softmax=False)
        hl, hr = S.raw_ctcnet_halo(net)
        plan = S.time_shard_plan(T, rank, world, hl, hr)
        x_ext = S.exchange_halo(x[:, :, plan["start"]:plan["end"]].contiguous(), plan, rank, world)
        assert torch.equal(x_ext, x[:, :, plan["lo"]:plan["hi"]])
        y = S.time_sharded_forward(lambda z: O.raw_ctcnet_forward(sd, z, layers, softmax=False), x_ext, plan, T,
                                   out_extra=net.feature_kwidth - 1)
        stop = plan["end"] + (2 if plan["end"] == T else 0)
        err_t = float((y - full[:, :, plan["start"]:stop]).abs().max())

        # ---------------- batch sharding: gradient all-reduce reproduces the single-process gradient --------
        torch.manual_seed(3)
        wl = [(6, 6, 2, d) for d in (1, 2)]
        wnet = W.WaveNet(6, 2, wl, 6, softmax=False)
        B, Tn = 4, 21
        lev = torch.randint(0, 6, (B, Tn))
        sig = torch.zeros(B, 6, Tn).scatter_(1, lev.unsqueeze(1), 1.0)

        def loss_of(sdict, s, scale_world):
            pred = O.wavenet_forward(sdict, s[:, :, 0:-1], wl, softmax=False)
            dense = torch.max(s[:, :, 1:], dim=1)[1]
            xe = O.xe_loss_sum_over_time(pred, dense)            # mean over the (local) batch
            ctc_like = pred.pow(2).sum()                         # any batch-SUM term
            return S.joint_loss_for_backward(xe, ctc_like, Tn, 7, scale_world)

        ref_sd = {k: v.detach().clone().requires_grad_(True) for k, v in wnet.state_dict().items()}
        loss_of(ref_sd, sig, 1).backward()
        s0, s1 = S.shard_range(B, rank, world)
        params = [torch.nn.Parameter(v.detach().clone()) for v in wnet.state_dict().values()]
        loc_sd = dict(zip(wnet.state_dict().keys(), params))
        loss_of(loc_sd, sig[s0:s1], world).backward()
        nb = S.allreduce_gradients(params, bucket_bytes=1024)
        err_g = max(float((p.grad - ref_sd[k].grad).abs().max()) for k, p in loc_sd.items() if p.grad is not None)
        q.put((rank, err_t, err_g, nb))
    finally:
        dist.destroy_process_group()


def test_time_and_batch_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_t, err_g, nb in res:
        assert err_t < 1e-5, (rank, err_t)          # time-sharded output == full-read output
        assert err_g < 1e-5, (rank, err_g)          # all-reduced gradient == single-process gradient
        assert nb >= 2                              # several buckets were exercised
