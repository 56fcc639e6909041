"""Per-call CUDA-event times of one config-2 WaveNet forward (the launches bench.py times)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wavenet_speech_b200 import _lib
w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
net, _ = bench.build_model(w)
net = net.cuda().bfloat16().eval()
x = bench.make_input(w, 0).cuda()
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    _lib.kernel_timing(True)
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    log = _lib.kernel_timing(False)
n = len(log) // 3
rows = [(name, sum(log[i + k * n][1].elapsed_time(log[i + k * n][2]) for k in range(3)) / 3) for i, (name, _a, _b) in enumerate(log[:n])]
tot = sum(r[1] for r in rows)
for name, ms in rows:
    print("%-44s %8.3f ms %5.1f%%" % (name, ms, 100 * ms / tot))
print("sum %.3f ms" % tot)
