"""Per-phase clock64 timeline of CTA 0 of the fused residual-block kernels (diagnostic).
usage: WNB200_TIMELINE=1 timeline.py <dilation> [v2|v3] [p[flags]]     (p = precise format; flags: experiment bits)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavenet_speech_b200 as W
from wavenet_speech_b200 import fastpath as FP

C, B, T = 256, 32, 16384
d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
toks = sys.argv[2:]
v2 = any(t in ("v2", "v3") for t in toks)
variant = 1 if "v2" in toks else (2 if "v3" in toks else 0)     # "v2" -> single-CTA, "v3" -> CTA pair
ptok = [t for t in toks if t.startswith("p")]
prec = bool(ptok)
xflags = int(ptok[0][1:] or 0) if prec else 0
nolo = "nolo" in toks
torch.manual_seed(0)
blk = W.ResidualBlock(C, C, 2, d, causal=True)
bn = torch.nn.Conv1d(C, C, 1)
pk = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in FP.pack_block(blk, bn, precise=prec).items()}
x = torch.randn(B, T, C, device="cuda").to(torch.float16 if prec else torch.bfloat16)
x_lo = (torch.randn(B, T, C, device="cuda") * 1e-3).half() if (prec and not nolo) else None
res_lo = torch.empty_like(x) if prec else None
variant |= xflags << 2
res = torch.empty_like(x)
skips = torch.zeros(B, T, C, device="cuda")
dbg = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(4):
    if i == 3:
        ev[0].record()
    if v2:
        FP.resblock(x, pk, res, skips, False, dbg=dbg, variant=variant, x_lo=x_lo, res_lo=res_lo)
    else:
        FP.chain(x, C, pk["offsets"], FP.TC_GATE, pk["w1"], pk["b1"], 2 * C, n2=2 * C, use_x2=1,
                 epi2=FP.EPI2_RESBLOCK, w2=pk["w2"], b2=pk["b2"], y_nlc=res, skips=skips, skips_init=0, dbg=dbg)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1])
print("kernel %.3f ms -> %.1f TFLOP/s (16*C*C flop/sample)" % (ms, 16 * C * C * B * T / ms / 1e9))
t = dbg.cpu().view(8, 16)
if v2:
    names = {0: "G1start", 1: "G1a_iss", 2: "G1b_iss", 3: "e1a_seen", 4: "e1b_seen", 5: "G2r_iss", 6: "G2s_iss",
             8: "E1a_go", 9: "E1b_go", 10: "E2a_go", 11: "E2b_go", 12: "tile_done", 13: "wfullG1a", 14: "wfullG1ab"}
else:
    names = {0: "G1start", 1: "G1_iss", 2: "act_ready", 3: "G2_iss", 4: "g1_full", 5: "act_done", 6: "g2_full",
             7: "tile_done"}
t0 = int(t[0, 0])
for it in range(6):
    print("tile", it, " ".join("%s=%d" % (n, int(t[it, i]) - (0 if n.startswith("wfull") else t0)) for i, n in sorted(names.items())))
