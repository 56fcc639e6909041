"""fp32 WaveNet (config-2 network) on the generic fp32-FMA kernels: the <= 1e-5 parity path's throughput.
    python scripts/bench_generic_fp32.py [B=4] [T=16384]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wavenet_speech_b200 as W

opts = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
B, T = int(opts.get("B", 4)), int(opts.get("T", 16384))
torch.manual_seed(0)
dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2
net = W.WaveNet(256, 2, [(256, 256, 2, d) for d in dil], 256, softmax=True).cuda().eval()
lev = torch.randint(0, 256, (B, T), device="cuda")
x = torch.zeros(B, 256, T, device="cuda").scatter_(1, lev.unsqueeze(1), 1.0)
with torch.no_grad():
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        net(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"what": "wavenet_config2_fp32_generic_kernels", "B": B, "T": T, "ms": round(ms, 3),
                  "samples_per_s": round(B * T / ms * 1e3), "tflops_as_written": round(21495808 * B * T / ms / 1e9, 1)}))
