"""
Golden fixtures for the ByteNet residual blocks and the incremental LinearConv1d path, written by the
REFERENCE's own modules (modules/block.py:86-173, modules/linear_conv_ops.py:39-68) imported from
/root/reference (authoring container only).  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_bytenet.py     # writes tests/golden/bytenet_*.npz, linearconv_stream_*.npz

Same file format as oracle/gen_golden.py (state_dict, inputs, outputs, meta).  The blocks' outputs come with
the gradients torch autograd computes through the reference module for the loss sum(y * probe).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("WN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
from modules.block import ResidualMUBlock, ResidualReLUBlock  # noqa: E402
from modules.linear_conv_ops import LinearConv1d  # noqa: E402


def save(name, module, inputs, outputs, meta, extra=None):
    blob = {}
    for k, v in module.state_dict().items():
        blob["sd/" + k] = v.detach().numpy()
    for k, v in inputs.items():
        blob["in/" + k] = v.detach().numpy()
    for k, v in outputs.items():
        blob["out/" + k] = v.detach().numpy()
    for k, v in (extra or {}).items():
        blob["extra/" + k] = np.asarray(v)
    blob["meta"] = np.array(json.dumps(meta))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print("%-32s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


def block_fixture(name, cls, nch, k, d, B, T, seed):
    torch.manual_seed(seed)
    m = cls(nch, k, d)
    m.init()
    with torch.no_grad():                       # gamma = 1 / beta = 0 after init(): move them so both are exercised
        for n, p in m.named_parameters():
            if n.endswith("gamma") or n.endswith("beta"):
                p.add_(torch.randn_like(p) * 0.2)
    x = torch.randn(B, nch, T, requires_grad=True)
    probe = torch.randn(B, nch, T)
    y = m(x)
    (y * probe).sum().backward()
    grads = {"grad/" + n: p.grad for n, p in m.named_parameters()}
    grads["grad_x"] = x.grad
    save(name, m, {"x": x.detach(), "probe": probe}, dict({"y": y.detach()}, **grads),
         {"nchannels": nch, "k": k, "d": d, "rf": int(m.receptive_field)})


def main():
    torch.set_num_threads(1)
    # tests/test_bytenet_blocks.py:9-13 shape, and a wider one (odd T, k = 3)
    block_fixture("bytenet_relu_block_c4", ResidualReLUBlock, 4, 2, 2, 3, 12, 811)
    block_fixture("bytenet_mu_block_c4", ResidualMUBlock, 4, 2, 2, 3, 12, 812)
    block_fixture("bytenet_relu_block_c24", ResidualReLUBlock, 24, 3, 2, 2, 37, 813)
    block_fixture("bytenet_mu_block_c24", ResidualMUBlock, 24, 3, 2, 2, 37, 814)

    # incremental evaluation: LinearConv1d.linear applied to every window of a sequence (what a decoder that
    # emits one frame at a time asks of it; tests/test_linear_conv_ops.py:10-15 shape and a wider one)
    for (name, cin, cout, k, d, B, T, seed) in [("linearconv_stream_k5_d3", 3, 3, 5, 3, 2, 15, 821),
                                                ("linearconv_stream_k3_d4", 20, 12, 3, 4, 3, 40, 822)]:
        torch.manual_seed(seed)
        m = LinearConv1d(cin, cout, k, dilation=d)
        rf = int(m.receptive_field)
        seq = torch.randn(B, cin, T)
        padded = torch.cat([torch.zeros(B, cin, rf - 1), seq], 2)        # causal start: zeros before the first frame
        ys = torch.stack([m.linear(padded[:, :, t:t + rf]) for t in range(T)], 2)
        save(name, m, {"seq": seq}, {"y": ys.detach()}, {"cin": cin, "cout": cout, "k": k, "d": d, "rf": rf})


if __name__ == "__main__":
    main()
