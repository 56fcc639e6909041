"""Time the gradient all-reduce alone (config-3 parameter set, 76 MB fp32) and report what NCCL uses."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import wavenet_speech_b200 as W
from wavenet_speech_b200 import sharding as S
rank, local, world = (int(os.environ.get(k, 0)) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2
wn = W.WaveNet(256, 2, [(256, 256, 2, d) for d in dil], 256, softmax=False).cuda()
cn = W.WaveNetClassifier(256, 5, [(256, 256, 2, d) for d in [1, 2, 4, 8, 16] * 3], 256, pool_kernel_size=3, softmax=False).cuda()
params = list(wn.parameters()) + list(cn.parameters())
for p in params:
    p.grad = torch.randn_like(p)
def timed(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {"world": world}
res["allreduce_gradients_ms"] = timed(lambda: S.allreduce_gradients(params))
flat = torch.randn(19_000_000, device="cuda")
res["flat_76MB_allreduce_ms"] = timed(lambda: dist.all_reduce(flat))
small = torch.randn(1024, device="cuda")
res["4KB_allreduce_ms"] = timed(lambda: dist.all_reduce(small))
res["n_params"] = len(params)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
