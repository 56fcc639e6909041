"""Device CTC loss (blank = 0, pre-softmax activations, summed over the batch: the warpctc_pytorch convention of
legacy_code/train.py:42-46) against the oracle's stand-in (torch CPU ctc_loss, pinned by the reference's own two
known answers 2.4628 and 1.4519: tests/test_classifier.py:53-59, ipynbs/CTC Overfit.ipynb cell 27)."""
import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import functional as WF
from oracle import wavenet_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu


def test_reference_known_answers():
    act = torch.tensor([[[.1, .6, .1, .1, .1]], [[.1, .1, .6, .1, .1]]])          # (T=2, B=1, C=5)
    loss = W.CTCLoss()(act.cuda(), torch.tensor([1, 2], dtype=torch.int32), torch.tensor([2]), torch.tensor([2]))
    assert loss.shape == (1,) and abs(float(loss) - 2.4628) < 1e-3
    act = torch.tensor([[[-10., -9., -8., -7., -6.]]])                             # (T=1, B=1, C=5)
    loss = W.CTCLoss()(act.cuda(), torch.tensor([3], dtype=torch.int32), torch.tensor([1]), torch.tensor([1]))
    assert abs(float(loss) - 1.4519) < 1e-3


@pytest.mark.parametrize("B,L,T,lens,layout,dtype", [
    (3, 5, 40, [7, 12, 0], "bct", torch.float32),
    (2, 5, 300, [100, 61], "tbc", torch.float32),
    (4, 8, 1500, [700, 300, 512, 1], "bct", torch.float32),        # > 1024 states: several states per thread
    (2, 5, 200, [60, 40], "bct", torch.bfloat16),
    (2, 40, 64, [20, 9], "tbc", torch.float32),
    (3, 5, 400, [150, 0, 1], "bct", torch.float32),                # an empty and a 1-label read beside a long one
    (2, 5, 200, [129, 64], "bct", torch.float32),
    (40, 5, 300, [90] * 40, "bct", torch.float32),                 # more reads than fit one wave of the grad kernel's first blocks
])
def test_ctc_matches_oracle(B, L, T, lens, layout, dtype):
    torch.manual_seed(B * 100 + T)
    act = (torch.randn(B, L, T) * 2).to(dtype).float()
    labels = torch.cat([torch.randint(1, L, (n,)) for n in lens]).int()
    if len(labels) > 3:
        labels[1] = labels[0]                                    # a repeated label (needs a blank in between)
    lengths = torch.tensor(lens, dtype=torch.int32)
    act_lens = torch.tensor([T - 3 * i for i in range(B)], dtype=torch.int32)
    # yardstick: the oracle in float64 (an fp32 log-domain CTC -- warp-ctc, torch -- is itself up to 1e-2 off at
    # T = 1500; the oracle in fp32 is held beside it to show the kernel is at least as close)
    a = act.double().requires_grad_(True)
    ref = O.ctc_loss_sum(a.permute(2, 0, 1), labels, act_lens, lengths, dtype=torch.float64)
    (ref * 0.5).backward()
    a32 = act.clone().requires_grad_(True)
    (O.ctc_loss_sum(a32.permute(2, 0, 1), labels, act_lens, lengths) * 0.5).backward()
    err32 = G.rel_linf(a32.grad, a.grad)
    g = (act.permute(2, 0, 1).contiguous() if layout == "tbc" else act).cuda().to(dtype).requires_grad_(True)
    loss = WF.ctc_loss_sum(g, labels.cuda(), lengths, act_lens, layout=layout)
    (loss * 0.5).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref))
    grad = g.grad.float().cpu()
    if layout == "tbc":
        grad = grad.permute(1, 2, 0)
    tol = 3e-4 if dtype == torch.float32 else 1e-2
    err = G.rel_linf(grad, a.grad)
    assert err <= tol and (dtype != torch.float32 or err <= max(err32, 2e-5)), (err, err32)


def test_ctc_infeasible_and_full_length_default():
    act = torch.randn(2, 5, 6)
    labels = torch.tensor([1, 1, 1, 1, 2, 3], dtype=torch.int32)       # read 0 needs 7 frames > 6: infeasible
    g = act.cuda().requires_grad_(True)
    loss = WF.ctc_loss_sum(g, labels.cuda(), torch.tensor([4, 2]))
    assert torch.isinf(loss)
    loss.backward()
    assert float(g.grad[0].abs().max()) == 0.0 and torch.isfinite(g.grad).all()
    ref = O.ctc_loss_sum(act[1:].permute(2, 0, 1), labels[4:], torch.tensor([6], dtype=torch.int32),
                         torch.tensor([2], dtype=torch.int32))
    a = act[1:].clone().requires_grad_(True)
    O.ctc_loss_sum(a.permute(2, 0, 1), labels[4:], torch.tensor([6], dtype=torch.int32),
                   torch.tensor([2], dtype=torch.int32)).backward()
    assert G.rel_linf(g.grad[1:].cpu(), a.grad) <= 1e-4 and float(ref) > 0


@pytest.mark.parametrize("B,L,T,dtype", [(3, 5, 50, torch.float32), (2, 5, 3000, torch.float32), (4, 8, 1025, torch.bfloat16),
                                         (1, 5, 1, torch.float32)])
def test_greedy_decode_matches_oracle(B, L, T, dtype):
    """Device argmax / collapse / drop-blank vs modules/sequence_decoders.py:9-23 + the notebook's collapse."""
    torch.manual_seed(T)
    # piecewise-constant winners so that repeats, blanks and chunk boundaries (1024 frames) all occur
    win = torch.randint(0, L, (B, (T + 2) // 3)).repeat_interleave(3, dim=1)[:, :T]
    x = (torch.randn(B, L, T) * 0.3).scatter_add_(1, win.unsqueeze(1), torch.full((B, 1, T), 3.0)).to(dtype)
    lens = torch.tensor([T - (7 * b) % max(1, T) for b in range(B)], dtype=torch.int32)
    ref_frames = O.argmax_decode(x.float().permute(0, 2, 1))
    got_frames = W.argmax_decode(x.cuda().permute(0, 2, 1))
    assert got_frames.dtype == torch.int64 and torch.equal(got_frames.cpu(), ref_frames)
    lab, n = W.ops.ctc_greedy_decode(x.cuda(), lens)
    for b in range(B):
        ref = O.collapse_decode(ref_frames[b, :int(lens[b])])
        assert int(n[b]) == len(ref) and lab[b, :int(n[b])].cpu().tolist() == [int(v) for v in ref], b
    strings = W.greedy_ctc_decode(x[:, :5].cuda())
    full = [O.collapse_decode(O.argmax_decode(x[:, :5].float().permute(0, 2, 1))[b]) for b in range(B)]
    assert strings == ["".join("_AGCT"[int(v)] for v in r) for r in full]
    assert W.Decoder('argmax').decode(x[:, :5].cuda())[1] == W.labels2strings(ref_frames.clamp(max=4)) or L > 5
