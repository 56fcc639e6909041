"""The oracle must reproduce the reference's own outputs (tests/golden/, written by
oracle/gen_golden.py from the reference nn.Modules) -- this is what pins it."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wavenet_oracle as O
from tests import _golden as G

torch.set_num_threads(1)


def _exact(a, b):
    # Same ATen CPU kernels in the same order => bit-identical; allow 2 ulp-ish slack for
    # platform differences in vectorised transcendental code paths.
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())


@pytest.mark.parametrize("name", [n for n in G.names() if "conv_k" in n and "linear" not in n])
def test_conv_ops(name):
    g = G.load(name)
    m = g["meta"]
    fn = O.causal_conv1d if m["causal"] else O.noncausal_conv1d
    y = fn(g["inp"]["x"], g["sd"]["conv1d.weight"], g["sd"]["conv1d.bias"], m["d"])
    _exact(y, g["out"]["y"])
    # independent tap-sum restatement pins the offsets
    y2 = O.conv1d_taps_numpy(g["inp"]["x"].numpy(), g["sd"]["conv1d.weight"].numpy(),
                             g["sd"]["conv1d.bias"].numpy(), O.tap_offsets(m["k"], m["d"], m["causal"]))
    np.testing.assert_allclose(y2, g["out"]["y"].numpy(), rtol=1e-4, atol=1e-5)


def test_tap_offsets_match_survey():
    assert O.tap_offsets(2, 1, False) == [-1, 0]
    assert O.tap_offsets(2, 2, False) == [-1, 1]
    assert O.tap_offsets(2, 4, False) == [-2, 2]
    assert O.tap_offsets(2, 16, False) == [-8, 8]
    assert O.tap_offsets(2, 3, False) == [-2, 1]
    assert O.tap_offsets(2, 512, True) == [-512, 0]
    assert O.tap_offsets(5, 3, True) == [-12, -9, -6, -3, 0]


@pytest.mark.parametrize("name", ["block_causal", "block_noncausal", "block_noncausal_k3_d3"])
def test_residual_block(name):
    g = G.load(name)
    m = g["meta"]
    res, skip = O.residual_block(g["sd"], "", g["inp"]["x"], m["d"], m["causal"])
    _exact(res, g["out"]["res"])
    _exact(skip, g["out"]["skip"])


@pytest.mark.parametrize("name", ["wavenet_test_shape", "wavenet_onehot_c32", "wavenet_c128_bf16w"])
def test_wavenet(name):
    g = G.load(name)
    m = g["meta"]
    y = O.wavenet_forward(g["sd"], g["inp"]["x"], m["layers"], softmax=m["softmax"])
    _exact(y, g["out"]["y"])
    if name == "wavenet_test_shape":   # reference tests/test_wavenet.py:34
        assert tuple(y.shape) == (5, 11, 14)


@pytest.mark.parametrize("name", ["rawctcnet_default", "rawctcnet_positions", "rawctcnet_causal",
                                  "rawctcnet_example_json", "rawctcnet_c128_bf16w"])
def test_raw_ctcnet(name):
    g = G.load(name)
    m = g["meta"]
    y = O.raw_ctcnet_forward(g["sd"], g["inp"]["x"], m["layers"], positions=m["positions"],
                             softmax=m["softmax"], causal=m["causal"])
    _exact(y, g["out"]["y"])
    assert y.shape[2] == g["inp"]["x"].shape[2] + m["feature_kwidth"] - 1


@pytest.mark.parametrize("name", ["classifier_pool3", "classifier_c128_bf16w"])
def test_classifier(name):
    g = G.load(name)
    m = g["meta"]
    y = O.classifier_forward(g["sd"], g["inp"]["x"], m["layers"], pool_kernel_size=m["pool_kernel_size"],
                             softmax=m["softmax"])
    _exact(y, g["out"]["y"])
    assert y.shape[2] == g["inp"]["x"].shape[2] // m["pool_kernel_size"]


def test_layernorm_linearconv_mu():
    g = G.load("layernorm_c6")
    _exact(O.layernorm(g["inp"]["x"], g["sd"]["gamma"], g["sd"]["beta"]), g["out"]["y"])
    g = G.load("linearconv_k3_d2")
    _exact(O.linear_conv1d_linear(g["inp"]["frame"], g["sd"]["weight"], g["sd"]["bias"], g["meta"]["d"]),
           g["out"]["y"])
    g = G.load("multiplicative_unit")
    _exact(O.multiplicative_unit(g["sd"], "", g["inp"]["x"], g["meta"]["d"]), g["out"]["y"])


@pytest.mark.parametrize("name", [n for n in G.names() if n.startswith("bytenet_")])
def test_bytenet_blocks_and_their_gradients(name):
    """ResidualReLUBlock / ResidualMUBlock (block.py:86-173): outputs AND the gradients torch autograd wrote through the
    reference module, reproduced by autograd through the oracle."""
    g = G.load(name)
    fn = O.residual_mu_block if "_mu_" in name else O.residual_relu_block
    sd = {k: v.clone().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["inp"]["x"].clone().requires_grad_(True)
    y = fn(sd, "", x, g["meta"]["d"])
    _exact(y.detach(), g["out"]["y"])
    (y * g["inp"]["probe"]).sum().backward()
    assert torch.allclose(x.grad, g["out"]["grad_x"], rtol=1e-4, atol=1e-6)
    for k, v in sd.items():
        ref = g["out"]["grad/" + k]
        assert torch.allclose(v.grad, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()) + 1e-7), k


@pytest.mark.parametrize("name", [n for n in G.names() if n.startswith("linearconv_stream")])
def test_linearconv_stream(name):
    g = G.load(name)
    y = O.linear_conv1d_stream(g["inp"]["seq"], g["sd"]["weight"], g["sd"]["bias"], g["meta"]["d"])
    _exact(y, g["out"]["y"])
    # frame-at-a-time evaluation == the causal convolution (conv_ops.py:39-44) of the whole sequence
    # (another summation order: held to fp32 round-off, not bit for bit)
    yc = O.causal_conv1d(g["inp"]["seq"], g["sd"]["weight"], g["sd"]["bias"], g["meta"]["d"])
    assert torch.allclose(yc, g["out"]["y"], rtol=1e-5, atol=1e-6), float((yc - g["out"]["y"]).abs().max())


def test_ctc_known_answers():
    # reference tests/test_classifier.py:53-59 -> 2.4628 ; ipynbs/CTC Overfit.ipynb cell 27 -> 1.4519
    acts = torch.tensor([[[.1, .6, .1, .1, .1]], [[.1, .1, .6, .1, .1]]])
    v = O.ctc_loss_sum(acts, torch.tensor([1, 2], dtype=torch.int32), torch.tensor([2], dtype=torch.int32),
                       torch.tensor([2], dtype=torch.int32))
    assert abs(float(v) - 2.4628) < 1e-3
    acts = torch.tensor([[[-10., -9., -8., -7., -6.]]])
    v = O.ctc_loss_sum(acts, torch.tensor([3], dtype=torch.int32), torch.tensor([1], dtype=torch.int32),
                       torch.tensor([1], dtype=torch.int32))
    assert abs(float(v) - 1.4519) < 1e-3


def test_train_step_losses_and_grads():
    g = G.load("train_step_small")
    m = g["meta"]
    wsd = {k: v.clone().requires_grad_(True) for k, v in g["other"]["wsd"].items()}
    csd = {k: v.clone().requires_grad_(True) for k, v in g["other"]["csd"].items()}
    wl = [tuple(l) for l in m["wave_layers"]]
    cl = [tuple(l) for l in m["cls_layers"]]
    avg_xe, avg_ctc, joint, pred, trans = O.train_step_losses(
        wsd, csd, g["inp"]["sig"], g["inp"]["seq"], g["inp"]["lengths"], wl, cl, m["pool"])
    _exact(pred.detach(), g["out"]["pred"])
    _exact(trans.detach(), g["out"]["trans"])
    T = g["inp"]["sig"].shape[2]
    assert abs(float(avg_xe) * T - float(g["out"]["xe"])) < 1e-3 * abs(float(g["out"]["xe"]))
    assert abs(float(joint) - float(g["out"]["joint"])) < 1e-4 * abs(float(g["out"]["joint"]))
    joint.backward()
    for k, ref in g["other"]["wgrad"].items():
        assert torch.allclose(wsd[k].grad, ref, rtol=1e-3, atol=1e-6), k
    for k, ref in g["other"]["cgrad"].items():
        assert torch.allclose(csd[k].grad, ref, rtol=1e-3, atol=1e-6), k


def test_flop_model_matches_baseline_md():
    layers = [(256, 256, 2, d) for d in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 2]
    assert O.block_flops(256, 256, 2, 256) == 1048576
    assert O.wavenet_flops_per_timestep(256, 2, layers, 256) == 21495808


def test_oracle_gradients_match_reference_written_gradients():
    """tests/golden/wavenet_c128_bf16w_grads.npz holds gradients computed by the REFERENCE's modules + torch autograd
    (oracle/gen_golden_tc.py).  Autograd through the oracle must reproduce them: this pins the backward half of the oracle
    that the tensor-core training tests lean on."""
    import torch
    from oracle import wavenet_oracle as O
    from tests import _golden as G
    g = G.load("wavenet_c128_bf16w_grads")
    layers = g["meta"]["layers"]
    sd = {k: v.clone().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["inp"]["x"].clone().requires_grad_(True)
    y = O.wavenet_forward(sd, x, layers, softmax=False)
    assert G.rel_linf(y.detach(), g["out"]["y"]) <= 2e-6
    (y * g["inp"]["R"]).sum().backward()
    refs = {k[len("grad/"):]: v for k, v in g["out"].items() if k.startswith("grad/")}
    assert G.rel_linf(x.grad, refs.pop("__input__")) <= 1e-5
    assert len(refs) > 20
    for n, ref in refs.items():
        assert G.rel_linf(sd[n].grad, ref) <= 1e-5, n
