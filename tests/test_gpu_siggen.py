"""On-device synthetic pore-model generator (csrc/siggen.cu) against the numpy restatement of the reference's
generator (utils/signal_gen.py <- utils/raw_signal_generator.py:77-118,189-203, utils/pore_model.py:58-96): the
deterministic part exactly on the same random draws, the draws themselves statistically."""
import numpy as np
import pytest
import torch

from wavenet_speech_b200.utils import signal_gen as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T", [(4, 4000), (2, 16384), (3, 37)])
def test_device_generator_matches_numpy_on_the_same_draws(B, T):
    sig, labels, (bases, reps, z) = S.device_raw_batch(B, T, seed=42 + T, return_draws=True)
    torch.cuda.synchronize()
    means, stdvs = S.load_pore_model()
    bases, reps, z, sig = bases.cpu().numpy(), reps.cpu().numpy(), z.cpu().numpy(), sig.cpu().numpy()[:, 0]
    assert bases.min() >= 1 and bases.max() <= 4 and reps.min() >= 1
    for b in range(B):
        kmers = S.kmer_indices(bases[b])
        seq = np.repeat(kmers, reps[b])[:T]
        assert len(seq) == T
        ref = means[seq] + stdvs[seq] * z[b]
        assert np.allclose(sig[b], ref, rtol=1e-6, atol=1e-4)
        n_used = int(np.searchsorted(np.cumsum(reps[b]), T, side="left")) + 1
        assert labels[b].cpu().numpy().tolist() == bases[b][2:2 + n_used].tolist()


def test_device_generator_statistics():
    sig, labels, (bases, reps, z) = S.device_raw_batch(8, 16384, seed=7, return_draws=True)
    bases, reps, z = bases.cpu().numpy(), reps.cpu().numpy().astype(np.float64), z.cpu().numpy().astype(np.float64)
    frac = np.bincount(bases.ravel(), minlength=5)[1:] / bases.size
    assert np.all(np.abs(frac - 0.25) < 0.01)
    # numpy's generator gives 2.95 +- 0.02 (tests/test_signal_gen_cpu.py); same law here
    rng = np.random.default_rng(0)
    ref = np.maximum((rng.gamma(S.DURATION_SHAPE, 1 / S.DURATION_RATE, size=200000) * S.SAMPLE_RATE).astype(int), 1)
    assert abs(reps.mean() - ref.mean()) < 0.05 and abs(reps.std() - ref.std()) < 0.08
    assert abs(np.mean(reps == 1) - np.mean(ref == 1)) < 0.01
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and abs((z ** 3).mean()) < 0.03
    assert abs((z ** 4).mean() - 3) < 0.1
    x = sig.cpu().numpy()
    assert 30 < x.min() and x.max() < 150
    # different seeds / reads give different signals; same seed reproduces
    again, _ = S.device_raw_batch(8, 16384, seed=7)
    assert torch.equal(again, sig) and not torch.equal(sig[0], sig[1])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_device_mu_law_one_hot(dtype):
    sig, _ = S.device_raw_batch(3, 5000, seed=3)
    oh, lev = S.device_one_hot(sig, 256, dtype=dtype, return_levels=True)
    torch.cuda.synchronize()
    x = sig.cpu().numpy()[:, 0]
    ref = np.stack([S.mu_law_levels(x[b], 256) for b in range(3)])
    got = lev.cpu().numpy()
    assert np.mean(got != ref) <= 1e-4 and np.abs(got - ref).max() <= 1      # bin-edge ties only
    assert oh.shape == (3, 256, 5000) and oh.dtype == dtype
    assert torch.equal(oh.float().argmax(1).cpu(), lev.cpu()) and torch.all(oh.float().sum(1) == 1)
