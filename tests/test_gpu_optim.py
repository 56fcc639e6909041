"""wavenet_speech_b200.optim.Adam (one launch over every parameter tensor) against torch.optim.Adam, the optimizer the
reference trains with (legacy_code/train.py:112-114, 55)."""
import copy

import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import _lib

pytestmark = pytest.mark.gpu


def _params(dtype, seed):
    torch.manual_seed(seed)
    shapes = [(256, 256, 2), (256,), (7,), (1, 5, 1), (100000,), (33, 129), (1,), (16384,), (16385,)]
    return [torch.randn(s).to(dtype).cuda() for s in shapes]


@pytest.mark.parametrize("dtype,wd", [(torch.float32, 0.0), (torch.float32, 1e-2), (torch.bfloat16, 0.0)])
def test_matches_torch_adam(dtype, wd):
    ours = [torch.nn.Parameter(p.clone()) for p in _params(dtype, 1)]
    ref = [torch.nn.Parameter(p.clone()) for p in _params(dtype, 1)]
    o1 = W.optim.Adam(ours, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    o2 = torch.optim.Adam(ref, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    for step in range(6):
        torch.manual_seed(100 + step)
        for a, b in zip(ours, ref):
            g = torch.randn(a.shape, device="cuda").to(dtype)
            a.grad = g.clone()                                            # fresh gradient tensors every step
            b.grad = g.clone()
        if step == 3:
            ours[2].grad = None                                           # a parameter that sits a step out keeps its own
            ref[2].grad = None                                            # step count (bias correction is per tensor)
        n0 = _lib.launch_count
        o1.step()
        assert _lib.launch_count - n0 == (1 if step <= 3 else 2)
        o2.step()
    tol = 2e-6 if dtype == torch.float32 else 1.6e-2
    for a, b in zip(ours, ref):
        err = float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))
        assert err <= tol, (tuple(a.shape), err)
    if dtype == torch.float32:
        for a, b in zip(ours, ref):
            sa, sb = o1.state[a], o2.state[b]
            assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-5, atol=1e-7)
            assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_state_dict_round_trip_with_torch():
    """State written by torch.optim.Adam continues under ours (and the other way round) with the same trajectory."""
    ps = [torch.nn.Parameter(p.clone()) for p in _params(torch.float32, 2)[:4]]
    pr = [torch.nn.Parameter(p.clone()) for p in _params(torch.float32, 2)[:4]]
    t = torch.optim.Adam(pr, lr=1e-2)
    for step in range(3):
        torch.manual_seed(step)
        for b in pr:
            b.grad = torch.randn_like(b)
        t.step()
    with torch.no_grad():
        for a, b in zip(ps, pr):
            a.copy_(b)
    o = W.optim.Adam(ps, lr=1e-2)
    o.load_state_dict(copy.deepcopy(t.state_dict()))      # (load_state_dict keeps the tensors it is handed)
    for step in range(3, 6):
        torch.manual_seed(step)
        for a, b in zip(ps, pr):
            g = torch.randn_like(b)
            a.grad, b.grad = g.clone(), g.clone()
        o.step()
        t.step()
    for a, b in zip(ps, pr):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    t2 = torch.optim.Adam(pr, lr=1e-2)
    t2.load_state_dict(copy.deepcopy(o.state_dict()))
    assert float(t2.state[pr[0]]["step"]) == 6.0
    for step in range(6, 8):
        torch.manual_seed(step)
        for a, b in zip(ps, pr):
            g = torch.randn_like(b)
            a.grad, b.grad = g.clone(), g.clone()
        o.step()
        t2.step()
    for a, b in zip(ps, pr):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
