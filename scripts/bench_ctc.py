"""CTC loss micro-benchmark at the config-3 shape (32 reads, 5461 frames, 2730 labels, 5 classes) and the RawCTCNet
training shape (128 reads, 4002 frames, ~1360 labels)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_speech_b200 import functional as WF, _lib


def run(B, T, n):
    torch.manual_seed(0)
    x = torch.randn(B, 5, T, device="cuda").bfloat16().requires_grad_(True)
    lengths = torch.full((B,), n, dtype=torch.int32)
    seq = torch.randint(1, 5, (B * n,), dtype=torch.int32, device="cuda")
    def step():
        x.grad = None
        WF.ctc_loss_sum(x, seq, lengths).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.kernel_timing(True)
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1 in _lib.kernel_timing(False):
        agg.setdefault(name, []).append(e0.elapsed_time(e1))
    print(json.dumps({"B": B, "T": T, "labels": n, **{k: round(sorted(v)[len(v) // 2], 3) for k, v in agg.items()}}))


run(32, 5461, 2730)
run(128, 4002, 1360)
