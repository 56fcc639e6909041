"""Channel-reduction kernels on rows that are not 16-byte multiples (the train step's T - 1 = 16383 frames,
legacy_code/train.py:30-39): the register-tile kernels with aligned loads and shifted frames (column_ops.cu, RegTileU)
against torch on the same values -- log-softmax + NLL forward, its backward, per-frame argmax -- for every phase of the
row start (T mod 8), ragged last tiles, channel counts below / at the tile's 256, fp32 and bf16 storage; and the
backward's element-wise boundary stores must not touch a byte outside the tensor."""
import pytest
import torch

import wavenet_speech_b200 as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,C,T", [(2, 256, 16383 // 8), (3, 256, 57), (2, 256, 113), (1, 200, 1001), (2, 128, 339),
                                   (2, 64, 55), (1, 40, 7), (2, 256, 2046), (1, 256, 1), (3, 5, 203)])
def test_xent_and_argmax_on_unaligned_rows(dtype, B, C, T):
    torch.manual_seed(B * 1000 + C + T)
    assert (T * (2 if dtype == torch.bfloat16 else 4)) % 16 != 0
    x = (torch.randn(B, C, T) * 3).to(dtype).cuda()
    tgt = torch.randint(0, C, (B, T), device="cuda")
    xf = x.float()
    # forward
    loss, lse = W.ops.xent_fwd(x, tgt)
    ref_lse = torch.logsumexp(xf, 1)
    ref_loss = ref_lse - xf.gather(1, tgt.unsqueeze(1)).squeeze(1)
    tol = 2e-3 if dtype == torch.bfloat16 else 2e-5          # bf16 storage: ex2.approx inside the sum
    assert float((lse - ref_lse).abs().max()) <= tol * max(1.0, float(ref_lse.abs().max()))
    assert float((loss - ref_loss).abs().max()) <= tol * max(1.0, float(ref_loss.abs().max()))
    # argmax (lowest index among ties, like torch)
    am = W.ops.argmax_channels(x)
    assert torch.equal(am, xf.argmax(1))
    # backward into a canary-framed buffer: (softmax - onehot) * gscale, nothing written outside
    gs = torch.full((1,), 0.37, device="cuda")
    n = x.numel()
    buf = torch.full((n + 64,), 7.0, dtype=dtype, device="cuda")
    dx = buf[32:32 + n].view(B, C, T)                        # 16-byte aligned like x, canaries either side
    W._lib.call("wnb200_xent_bwd", W.ops._dt(x), B, C, T, W.ops._p(x), W.ops._p(tgt), W.ops._p(ref_lse.contiguous()),
                W.ops._p(gs), W.ops._p(dx), W.ops._stream())
    assert float((buf[:32].float() - 7.0).abs().max()) == 0.0 and float((buf[32 + n:].float() - 7.0).abs().max()) == 0.0
    ref_dx = (torch.softmax(xf, 1) - torch.nn.functional.one_hot(tgt, C).permute(0, 2, 1).float()) * 0.37
    btol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert float((dx.float() - ref_dx).abs().max()) <= btol * max(1e-3, float(ref_dx.abs().max()))


def test_ties_and_extremes_on_unaligned_rows():
    """All-equal columns (argmax 0), -inf-like entries, and a tensor whose last vector ends exactly at the allocation."""
    B, C, T = 2, 256, 59
    x = torch.zeros(B, C, T, device="cuda").bfloat16()
    assert int(W.ops.argmax_channels(x).abs().max()) == 0
    x[:, 17, :] = 5.0
    x[:, 200, 10:20] = 5.0                                   # tie: the lower channel wins
    x[:, 3, 30] = 9.0
    am = W.ops.argmax_channels(x)
    ref = x.float().argmax(1)
    assert torch.equal(am, ref)
    x[:, 100:, :] = -30000.0
    loss, lse = W.ops.xent_fwd(x, torch.full((B, T), 17, device="cuda", dtype=torch.int64))
    assert torch.isfinite(loss).all() and torch.isfinite(lse).all()
    assert float((lse - torch.logsumexp(x.float(), 1)).abs().max()) <= 2e-3 * 10
