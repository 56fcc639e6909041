"""Host-side logic of the second session's additions, checked without a GPU: the MultiplicativeUnit weight packing
against the row order include/wnb200.h documents for WNB200_EPI_MU, the tap-major weight matrices of the tensor-core
ByteNet form, eligibility / argument checks that must refuse before any kernel is called, and that CPU tensors raise."""
import numpy as np
import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import functional as WF


def test_mu_weight_packing_follows_the_header_row_order():
    """wnb200_taps_t.mu_h: per 32 channels, rows [0, 64) = channel r / 4 of the group, unit r % 4 in the order
    (gate1, gate2, gate3, update); rows [64, 128) = channel 16 + (r - 64) / 4 likewise; channels past M are zero rows."""
    torch.manual_seed(0)
    for (M, C, k) in [(6, 6, 3), (32, 32, 2), (33, 33, 1), (80, 80, 4)]:
        mu = W.MultiplicativeUnit(M, k, dilation=2)
        wk, bk = WF._pack_mu(mu.convs, torch.float32)
        nt = (M + 31) // 32
        assert tuple(wk.shape) == (k, nt * 128, C) and tuple(bk.shape) == (nt * 128,)
        units = [c.weight.detach().numpy() for c in mu.convs]
        biases = [c.bias.detach().numpy() for c in mu.convs]
        for R in range(nt * 128):
            tile, r = divmod(R, 128)
            ch = tile * 32 + (r // 4 if r < 64 else 16 + (r - 64) // 4)
            u = r % 4
            for j in range(k):
                want = units[u][ch, :, j] if ch < M else np.zeros(C, dtype=np.float32)
                np.testing.assert_array_equal(wk[j, R].numpy(), want)
            assert float(bk[R]) == (float(biases[u][ch]) if ch < M else 0.0)


def test_tensor_core_weight_matrix_is_tap_major():
    from wavenet_speech_b200 import bytenet_tc
    torch.manual_seed(1)
    w = torch.randn(8, 5, 3)
    m = bytenet_tc._wmat(w)
    assert m.dtype == torch.bfloat16 and tuple(m.shape) == (8, 15)
    for j in range(3):
        assert torch.equal(m[:, j * 5:(j + 1) * 5], w[:, :, j].to(torch.bfloat16))      # wnb200_dense_t.w: [N][ntaps * Cin]
    m2 = bytenet_tc._wmat(torch.randn(4, 7))                                           # nn.Linear-shaped weight
    assert tuple(m2.shape) == (4, 7)


def test_tensor_core_form_is_refused_off_its_domain():
    from wavenet_speech_b200 import bytenet_tc
    blk = W.ResidualReLUBlock(256, 2, 2)
    with torch.no_grad():
        assert not bytenet_tc.eligible(blk, torch.zeros(2, 256, 64))                         # CPU tensor
        assert not bytenet_tc.eligible(blk, torch.zeros(2, 256, 64, dtype=torch.bfloat16))   # still CPU
        assert not bytenet_tc.eligible(W.ResidualReLUBlock(96, 2, 2), torch.zeros(1, 96, 8)) # C not a multiple of 128
        assert not bytenet_tc.eligible(W.ResidualReLUBlock(256, 5, 1), torch.zeros(1, 256, 8))


def test_cpu_tensors_raise_on_the_new_entry_points():
    for cls in (W.ResidualReLUBlock, W.ResidualMUBlock):
        with pytest.raises(RuntimeError, match="CUDA"):
            with torch.no_grad():
                cls(8, 2, 2)(torch.zeros(1, 8, 16))
    conv = W.LinearConv1d(4, 6, 3, dilation=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            conv.linear(torch.zeros(2, 4, conv.receptive_field))
    with pytest.raises(NotImplementedError):
        W.modules.linear_conv_ops.ConvStream(W.LinearConv1d(4, 6, 3, padding=1), 2, torch.float32, "cpu")
    with pytest.raises(NotImplementedError):
        W.optim.Adam([torch.nn.Parameter(torch.zeros(3))], amsgrad=True)
    with pytest.raises(ValueError):
        W.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=-1.0)
    opt = W.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=1e-3)
    opt.param_groups[0]["params"][0].grad = torch.zeros(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()


def test_bytenet_blocks_keep_the_reference_layout():
    """Same constructor arguments, stack indices, state_dict keys and receptive field as the reference's blocks
    (block.py:86-173): the fixtures written by the reference load with strict=True."""
    from tests import _golden as G
    for name in [n for n in G.names() if n.startswith("bytenet_")]:
        g = G.load(name)
        cls = W.ResidualMUBlock if "_mu_" in name else W.ResidualReLUBlock
        blk = cls(g["meta"]["nchannels"], g["meta"]["k"], g["meta"]["d"])
        blk.load_state_dict(g["sd"], strict=True)
        assert blk.receptive_field == g["meta"]["rf"]
        assert sorted(blk.state_dict().keys()) == sorted(g["sd"].keys())
    for name in [n for n in G.names() if n.startswith("linearconv_stream")]:
        g = G.load(name)
        m = g["meta"]
        conv = W.LinearConv1d(m["cin"], m["cout"], m["k"], dilation=m["d"])
        conv.load_state_dict(g["sd"], strict=True)
        assert conv.receptive_field == m["rf"] and conv._ker_ixs == list(range(0, m["rf"], m["d"]))
