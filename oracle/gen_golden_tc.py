"""
Extra golden fixtures at shapes the TENSOR-CORE path accepts (128 channels), written by the REFERENCE's own modules
(imported from /root/reference: authoring container only).  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_tc.py         # adds tests/golden/*_c128_bf16w.npz, leaves the other fixtures alone

The parameters and inputs are rounded to bf16-representable values BEFORE the reference runs (in fp32), so the stored
output is the reference's answer for exactly the numbers the tensor-core kernels see: the `-m gpu` tests compare the
tcgen05 path with the reference directly (<= 2e-2 on logits, the stated bf16 tolerance), not only through the oracle.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402  (imports the reference modules, provides save / randomize_biases)


def _round_params_to_bf16(m):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.bfloat16().float())


def main():
    torch.set_num_threads(1)
    C = 128
    layers = [(C, C, 2, 1), (C, C, 2, 2)]

    torch.manual_seed(1301)
    m = GG.WaveNet(C, 2, layers, C, softmax=False)
    GG.randomize_biases(m, 0.05)
    _round_params_to_bf16(m)
    lev = torch.randint(0, C, (2, 160))
    x = torch.zeros(2, C, 160).scatter_(1, lev.unsqueeze(1), 1.0)
    GG.save("wavenet_c128_bf16w", m, {"x": x}, {"y": m(x)},
            {"in_dim": C, "entry_kwidth": 2, "layers": layers, "out_dim": C, "softmax": False})

    torch.manual_seed(1302)
    m = GG.RawCTCNet(C, 3, 5, layers, C, softmax=False)
    GG.randomize_biases(m, 0.05)
    _round_params_to_bf16(m)
    x = torch.randn(2, 1, 150).bfloat16().float()
    GG.save("rawctcnet_c128_bf16w", m, {"x": x}, {"y": m(x)},
            {"num_features": C, "feature_kwidth": 3, "num_labels": 5, "layers": layers, "out_dim": C,
             "positions": False, "softmax": False, "causal": False})

    torch.manual_seed(1303)
    m = GG.WaveNetClassifier(C, 5, layers, C, pool_kernel_size=3, softmax=False)
    GG.randomize_biases(m, 0.05)
    _round_params_to_bf16(m)
    x = torch.randn(2, C, 151).bfloat16().float()
    GG.save("classifier_c128_bf16w", m, {"x": x}, {"y": m(x)},
            {"in_dim": C, "num_labels": 5, "layers": layers, "out_dim": C, "pool_kernel_size": 3, "softmax": False})

    # ---- reference-written GRADIENTS at a tensor-core shape: the reference's own modules and
    # torch autograd, parameters and input bf16-representable, loss = sum(y * R) for a stored R.  The tensor-core
    # training path is compared with these numbers directly -- an oracle nobody modified.
    torch.manual_seed(1304)
    glayers = [(C, C, 2, 1), (C, C, 2, 4)]
    m = GG.WaveNet(C, 2, glayers, C, softmax=False)
    GG.randomize_biases(m, 0.05)
    _round_params_to_bf16(m)
    lev = torch.randint(0, C, (2, 200))
    x = (torch.zeros(2, C, 200).scatter_(1, lev.unsqueeze(1), 1.0) + 0.05 * torch.randn(2, C, 200)).bfloat16().float()
    x.requires_grad_(True)
    R = torch.randn(2, C, 200).bfloat16().float()
    y = m(x)
    (y * R).sum().backward()
    grads = {"grad/" + n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    grads["grad/__input__"] = x.grad.detach().clone()
    GG.save("wavenet_c128_bf16w_grads", m, {"x": x.detach(), "R": R}, dict({"y": y.detach()}, **grads),
            {"in_dim": C, "entry_kwidth": 2, "layers": glayers, "out_dim": C, "softmax": False})


if __name__ == "__main__":
    main()
