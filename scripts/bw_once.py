"""Two launches of each second-session bandwidth kernel on the config-2 tensor (for `ncu --set full`, scripts/ncu_bw2.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wavenet_speech_b200 import fastpath as FP, ops, _lib
B, C, T = 32, 256, 16384
x = torch.randn(B, C, T, device="cuda", dtype=torch.bfloat16)
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
for _ in range(2):
    FP.ncl_to_nlc_bf16(x)
    y = torch.empty(B, T // 3, C, device="cuda", dtype=torch.bfloat16)
    _lib.call("wnb200_avgpool_ncl_to_nlc", 1, B, C, T, 3, ops._p(x), 0, ops._p(y), ops._stream())
    ops.softmax_fwd(x)
    ops.layernorm_fwd(x, g, b, 1e-6)
torch.cuda.synchronize()
