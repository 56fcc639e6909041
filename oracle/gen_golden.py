"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE's own
nn.Modules (imported from /root/reference, which only exists in the authoring
container) on seeded weights and inputs.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Each fixture stores the reference state_dict (so nothing depends on torch's RNG
stream), the input, the reference output, and a JSON `meta` blob with the
constructor arguments.  tests/test_oracle_golden.py replays them through
oracle/wavenet_oracle.py (must match bit-for-bit in fp32) and the `-m gpu` tests
replay them through the CUDA path.

It also exports the reference's 5-mer Gaussian pore-model table
(utils/r9.4_450bps.5mer.template.npz: 1024 means, 1024 stdvs) as the data file
the synthetic-signal generator reads.
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
from modules.block import ResidualBlock, MultiplicativeUnit  # noqa: E402
from modules.classifier import WaveNetClassifier  # noqa: E402
from modules.conv_ops import CausalConv1d, NonCausalConv1d  # noqa: E402
from modules.layernorm import LayerNorm  # noqa: E402
from modules.linear_conv_ops import LinearConv1d  # noqa: E402
from modules.raw_ctcnet import RawCTCNet  # noqa: E402
from modules.wavenet import WaveNet  # noqa: E402


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
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


def randomize_biases(module, scale=0.1):
    """The reference zero-initialises most biases (wavenet.py:74-85); perturb them so the
    fixtures exercise the bias path too."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * scale)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # --- conv ops (tests/test_conv_ops.py:10-36 shapes) --------------------------------------
    torch.manual_seed(101)
    m = CausalConv1d(4, 6, 5, dilation=3)
    x = torch.randn(3, 4, 15)
    save("causal_conv_k5_d3", m, {"x": x}, {"y": m(x)}, {"cin": 4, "cout": 6, "k": 5, "d": 3, "causal": True})
    for (k, d) in [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 5), (2, 16)]:
        torch.manual_seed(110 + 10 * k + d)
        m = NonCausalConv1d(4, 6, k, dilation=d)
        x = torch.randn(2, 4, 37)
        save("noncausal_conv_k%d_d%d" % (k, d), m, {"x": x}, {"y": m(x)},
             {"cin": 4, "cout": 6, "k": k, "d": d, "causal": False})

    # --- residual block (tests/test_block.py:9-40 shapes) ----------------------------------
    for causal in (True, False):
        torch.manual_seed(202)
        m = ResidualBlock(4, 5, 2, 2, causal=causal)
        randomize_biases(m)
        x = torch.randn(3, 4, 12)
        res, skip = m(x)
        save("block_%s" % ("causal" if causal else "noncausal"), m, {"x": x}, {"res": res, "skip": skip},
             {"cin": 4, "cout": 5, "k": 2, "d": 2, "causal": causal})
    torch.manual_seed(203)
    m = ResidualBlock(8, 8, 3, 3, causal=False)
    randomize_biases(m)
    x = torch.randn(2, 8, 40)
    res, skip = m(x)
    save("block_noncausal_k3_d3", m, {"x": x}, {"res": res, "skip": skip},
         {"cin": 8, "cout": 8, "k": 3, "d": 3, "causal": False})

    # --- WaveNet (tests/test_wavenet.py:10-34: seq_dim 11, 40 layers d=1..512 x4, B=5, T=14) ----
    torch.manual_seed(303)
    dil = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 4
    layers = [(11, 11, 2, d) for d in dil]
    m = WaveNet(11, 2, layers, 11, softmax=True)
    randomize_biases(m, 0.05)
    x = torch.randn(5, 11, 14)
    save("wavenet_test_shape", m, {"x": x}, {"y": m(x)},
         {"in_dim": 11, "entry_kwidth": 2, "layers": layers, "out_dim": 11, "softmax": True})

    torch.manual_seed(304)
    layers = [(32, 32, 2, d) for d in [1, 2, 4, 8, 16, 32]]
    m = WaveNet(32, 2, layers, 32, softmax=False)
    randomize_biases(m, 0.05)
    lev = torch.randint(0, 32, (2, 100))
    x = torch.zeros(2, 32, 100).scatter_(1, lev.unsqueeze(1), 1.0)      # one-hot, like the real input
    save("wavenet_onehot_c32", m, {"x": x}, {"y": m(x)},
         {"in_dim": 32, "entry_kwidth": 2, "layers": layers, "out_dim": 32, "softmax": False})

    # --- RawCTCNet --------------------------------------------------------------------------
    torch.manual_seed(404)
    layers = [(16, 16, 2, d) for d in [1, 2, 4, 8]] + [(16, 16, 3, 2)]
    for tag, kw in [("default", dict(softmax=False)),
                    ("positions", dict(softmax=True, positions=True)),
                    ("causal", dict(softmax=False, causal=True))]:
        m = RawCTCNet(16, 3, 5, layers, 16, **kw)
        x = torch.randn(2, 1, 50)
        meta = {"num_features": 16, "feature_kwidth": 3, "num_labels": 5, "layers": layers, "out_dim": 16,
                "softmax": kw.get("softmax", True), "positions": kw.get("positions", False),
                "causal": kw.get("causal", False)}
        save("rawctcnet_%s" % tag, m, {"x": x}, {"y": m(x)}, meta)

    # BASELINE config 1: RawCTCNet from configs/example.json (SURVEY 5.6 adapter), reduced B x T
    torch.manual_seed(405)
    layers = [(1, 1, 1, 1)]
    m = RawCTCNet(256, 2, 8, layers, 256, softmax=False)
    x = torch.randn(2, 1, 300)
    y = m(x)
    save("rawctcnet_example_json", m, {"x": x}, {"y": y},
         {"num_features": 256, "feature_kwidth": 2, "num_labels": 8, "layers": layers, "out_dim": 256,
          "softmax": False, "positions": False, "causal": False})

    # --- WaveNetClassifier (tests/test_classifier.py:9-40, reduced) ---------------------------
    torch.manual_seed(505)
    layers = [(16, 16, 2, d) for d in [1, 2, 4, 8, 16]]
    m = WaveNetClassifier(16, 5, layers, 16, pool_kernel_size=3, softmax=False)
    randomize_biases(m, 0.05)
    x = torch.randn(2, 16, 100)
    save("classifier_pool3", m, {"x": x}, {"y": m(x)},
         {"in_dim": 16, "num_labels": 5, "layers": layers, "out_dim": 16, "pool_kernel_size": 3,
          "softmax": False})

    # --- LayerNorm / LinearConv1d.linear / MultiplicativeUnit -------------------------------
    torch.manual_seed(606)
    m = LayerNorm(6)
    with torch.no_grad():
        m.gamma.add_(torch.randn_like(m.gamma) * 0.1)
        m.beta.add_(torch.randn_like(m.beta) * 0.1)
    x = torch.randn(2, 6, 9)
    save("layernorm_c6", m, {"x": x}, {"y": m(x)}, {"features": 6, "dim": 1, "eps": 1e-6})

    torch.manual_seed(607)
    m = LinearConv1d(4, 6, 3, dilation=2)
    frame = torch.randn(2, 4, m.receptive_field)
    save("linearconv_k3_d2", m, {"frame": frame}, {"y": m.linear(frame)},
         {"cin": 4, "cout": 6, "k": 3, "d": 2, "rf": int(m.receptive_field)})

    torch.manual_seed(608)
    m = MultiplicativeUnit(6, 3, dilation=2)
    x = torch.randn(2, 6, 20)
    save("multiplicative_unit", m, {"x": x}, {"y": m(x)}, {"ndim": 6, "k": 3, "d": 2})

    # --- train step (legacy_code/train.py:24-55) with torch's CTC standing in for warp-ctc ------
    torch.manual_seed(707)
    wl = [(16, 16, 2, d) for d in [1, 2, 4, 8]]
    cl = [(16, 16, 2, d) for d in [1, 2, 4]]
    wn = WaveNet(16, 2, wl, 16, softmax=False)
    cn = WaveNetClassifier(16, 5, cl, 16, pool_kernel_size=3, softmax=False)
    B, T = 3, 61
    lev = torch.randint(0, 16, (B, T))
    sig = torch.zeros(B, 16, T).scatter_(1, lev.unsqueeze(1), 1.0)
    lengths = torch.tensor([4, 6, 5], dtype=torch.int32)
    seq = torch.randint(1, 5, (int(lengths.sum()),), dtype=torch.int32)
    pred = wn(sig[:, :, 0:-1])
    trans = cn(pred)
    dense = torch.max(sig[:, :, 1:], dim=1)[1]
    xe = 0.
    for t in range(T - 1):
        xe = xe + torch.nn.CrossEntropyLoss()(pred[:, :, t], dense[:, t])
    probs = trans.permute(2, 0, 1).contiguous()
    pl = torch.full((B,), probs.shape[0], dtype=torch.int32)
    ctc = torch.nn.functional.ctc_loss(torch.log_softmax(probs, 2), seq, pl, lengths, blank=0, reduction="sum")
    joint = xe / sig.size(2) + ctc / trans.size(2)
    joint.backward()
    blob = {}
    for k, v in wn.state_dict().items():
        blob["wsd/" + k] = v.numpy()
    for k, v in cn.state_dict().items():
        blob["csd/" + k] = v.numpy()
    for n, p in wn.named_parameters():
        if n in ("entry_conv1d.conv1d.weight", "convolutions.1.conv_tanh.conv1d.weight",
                 "convolutions.2.residual_proj.weight", "bottlenecks.2.weight", "output_stack.3.bias"):
            blob["wgrad/" + n] = p.grad.numpy()
    for n, p in cn.named_parameters():
        if n in ("input_block.conv_sigmoid.conv1d.weight", "convolutions.2.conv1x1_skip.weight",
                 "output_block.1.weight"):
            blob["cgrad/" + n] = p.grad.numpy()
    blob.update({"in/sig": sig.numpy(), "in/seq": seq.numpy(), "in/lengths": lengths.numpy(),
                 "out/xe": xe.detach().numpy(), "out/ctc": ctc.detach().numpy(),
                 "out/joint": joint.detach().numpy(), "out/pred": pred.detach().numpy(),
                 "out/trans": trans.detach().numpy(),
                 "meta": np.array(json.dumps({"wave_layers": wl, "cls_layers": cl, "pool": 3, "dim": 16,
                                              "num_labels": 5}))})
    path = os.path.join(OUT, "train_step_small.npz")
    np.savez_compressed(path, **blob)
    print("%-28s %8.1f KB" % ("train_step_small", os.path.getsize(path) / 1024.0))

    # --- pore model table for the synthetic-signal generator -----------------------------------
    z = np.load(os.path.join(REF, "utils", "r9.4_450bps.5mer.template.npz"))
    table = np.stack([z["means"].astype(np.float32), z["stdvs"].astype(np.float32)], axis=0)
    dst = os.path.join(ROOT, "wavenet_speech_b200", "utils", "pore_model_r94_5mer.npy")
    np.save(dst, table)
    print("pore model table ->", dst, table.shape)


if __name__ == "__main__":
    main()
