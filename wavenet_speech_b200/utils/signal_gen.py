"""Synthetic Gaussian pore-model signal: the benchmark / test input of the hot path.

Restates, vectorised in numpy, what the reference's on-line generators do (paths relative to the reference):
  * k-mer index of a centred 5-base window, sum((nt-1) * [256,64,16,4,1]), edges dropped
    (utils/raw_signal_generator.py:91-93,107-108);
  * samples per k-mer ~ max(1, int(Gamma(shape=2.461964, scale=1/587.2858) * 800))
    (raw_signal_generator.py:77-78,189-203);
  * picoamp samples ~ N(means[k], stdvs[k]) from the r9.4 450bps 5-mer template table
    (raw_signal_generator.py:110-118), shipped here as pore_model_r94_5mer.npy (2 x 1024 float32);
  * bases ~ U{1..4} (utils/gaussian_kmer_model.py:281) in place of the reference-genome HDF5 the reference
    reads, which is not distributed with it;
  * mu-law 256-level quantisation and one-hot encoding (utils/pore_model.py:58-62,78-96).
Host-side input preparation only -- not part of the timed path."""
import os

import numpy as np

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pore_model_r94_5mer.npy")
DURATION_SHAPE, DURATION_RATE, SAMPLE_RATE = 2.461964, 587.2858, 800.0
_KMER_W = np.array([256, 64, 16, 4, 1], dtype=np.int64)


def load_pore_model():
    t = np.load(_TABLE)
    return t[0], t[1]


def kmer_indices(bases):
    """bases: int array of values 1..4, length L -> k-mer ids for the L-4 fully covered windows."""
    b = np.asarray(bases, dtype=np.int64) - 1
    win = np.lib.stride_tricks.sliding_window_view(b, 5)
    return win @ _KMER_W


def signal_from_draws(bases, reps, z, T=None):
    """The deterministic part of the generator: bases (1..4), samples per k-mer and standard-normal draws in, picoamp
    samples out (float64) -- what utils/raw_signal_generator.py:101-118 computes once its random draws are fixed.
    tests/golden/siggen_reference.npz holds the reference's own outputs for recorded draws."""
    means, stdvs = load_pore_model()
    seq = np.repeat(kmer_indices(bases), np.asarray(reps, dtype=np.int64))
    if T is not None:
        seq = seq[:T]
    z = np.asarray(z, dtype=np.float64)[:len(seq)]
    return means[seq].astype(np.float64) + stdvs[seq].astype(np.float64) * z


def raw_signal(T, rng, with_labels=False):
    """One read of exactly T picoamp samples (float32).  Optionally also returns the bases (labels 1..4)
    whose k-mers were (at least partly) emitted."""
    means, stdvs = load_pore_model()
    nb = int(T / 2.5) + 64
    while True:
        bases = rng.integers(1, 5, size=nb)
        kmers = kmer_indices(bases)
        reps = (rng.gamma(DURATION_SHAPE, 1.0 / DURATION_RATE, size=kmers.shape) * SAMPLE_RATE).astype(np.int64)
        reps = np.maximum(reps, 1)
        if int(reps.sum()) >= T:
            break
        nb *= 2
    seq = np.repeat(kmers, reps)[:T]
    sig = rng.normal(means[seq], stdvs[seq]).astype(np.float32)
    if not with_labels:
        return sig
    n_used = int(np.searchsorted(np.cumsum(reps), T, side="left")) + 1
    return sig, bases[2:2 + n_used].astype(np.int64)


def raw_batch(B, T, seed=1234, normalize=True, with_labels=False):
    """(B, 1, T) float32 raw signal; normalised over the batch the way the reference's callers do with
    BatchNorm1d(1) (legacy_code/run_raw_ctc.py:38,58)."""
    rng = np.random.default_rng(seed)
    sigs, labels = [], []
    for _ in range(B):
        r = raw_signal(T, rng, with_labels=with_labels)
        if with_labels:
            sigs.append(r[0])
            labels.append(r[1])
        else:
            sigs.append(r)
    x = np.stack(sigs, 0)[:, None, :]
    if normalize:
        x = (x - x.mean()) / (x.std() + 1e-5)
    x = x.astype(np.float32)
    return (x, labels) if with_labels else x


def mu_law_levels(sig, num_levels=256):
    """utils/pore_model.py:58-62,78-86: normalise by (max-min), mu-law compand with mu=num_levels, digitize on
    linspace(-1,1,num_levels).  Returns integer levels clipped into [0, num_levels-1]."""
    sig = np.asarray(sig, dtype=np.float64)
    norm = (sig - sig.mean()) / (sig.max() - sig.min())
    mu = float(num_levels)
    mapped = np.sign(norm) * np.log1p(mu * np.abs(norm)) / np.log1p(mu)
    lev = np.digitize(mapped, np.linspace(-1.0, 1.0, num=num_levels))
    return np.clip(lev, 0, num_levels - 1).astype(np.int64)


def quantized_batch(B, T, num_levels=256, seed=1234, with_labels=False):
    """(B, T) int64 levels of B synthetic reads (one-hot encode with `one_hot`)."""
    rng = np.random.default_rng(seed)
    levs, labels = [], []
    for _ in range(B):
        r = raw_signal(T, rng, with_labels=with_labels)
        if with_labels:
            levs.append(mu_law_levels(r[0], num_levels))
            labels.append(r[1])
        else:
            levs.append(mu_law_levels(r, num_levels))
    levs = np.stack(levs, 0)
    return (levs, labels) if with_labels else levs


def one_hot(levels, num_levels=256, dtype=np.float32):
    """(B, T) levels -> (B, num_levels, T) one-hot (utils/pore_model.py:88-96)."""
    B, T = levels.shape
    out = np.zeros((B, num_levels, T), dtype=dtype)
    out[np.arange(B)[:, None], levels, np.arange(T)[None, :]] = 1
    return out


# ------------------------------------------------------------------------------------------------ on the device
def device_raw_batch(B, T, seed=1234, device="cuda", return_draws=False):
    """Same generator on the GPU (csrc/siggen.cu): -> (sig fp32 (B, 1, T) in pA, labels: list of int64 tensors with
    values 1..4).  With return_draws also the random draws (bases, reps, z) for an exact check against the numpy
    restatement above."""
    import torch

    from .. import _lib, ops
    ops.check_device()
    means, stdvs = load_pore_model()
    dev = torch.device(device)
    m, s = torch.from_numpy(means.copy()).to(dev), torch.from_numpy(stdvs.copy()).to(dev)
    nb = int(T / 2.5) + 64
    while True:
        bases = torch.empty((B, nb), dtype=torch.int32, device=dev)
        reps = torch.empty((B, nb - 4), dtype=torch.int32, device=dev)
        n_used = torch.empty((B,), dtype=torch.int32, device=dev)
        z = torch.empty((B, T), dtype=torch.float32, device=dev) if return_draws else None
        sig = torch.empty((B, T), dtype=torch.float32, device=dev)
        _lib.call("wnb200_siggen_raw", B, T, nb, int(seed), ops._p(m), ops._p(s), ops._p(bases), ops._p(reps),
                  ops._p(n_used), ops._p(z), ops._p(sig), ops._stream())
        nu = n_used.cpu()
        if int(nu.min()) >= 0:
            break
        nb *= 2
    labels = [bases[b, 2:2 + int(nu[b])].long() for b in range(B)]
    if return_draws:
        return sig.unsqueeze(1), labels, (bases, reps, z)
    return sig.unsqueeze(1), labels


def device_one_hot(sig, num_levels=256, dtype=None, return_levels=False):
    """(B, 1, T) or (B, T) fp32 pA signal on the GPU -> one-hot (B, num_levels, T) of its per-read mu-law levels."""
    import torch

    from .. import _lib, ops
    x = sig.reshape(sig.shape[0], -1).float().contiguous()
    B, T = x.shape
    dtype = dtype or torch.bfloat16
    oh = torch.empty((B, num_levels, T), dtype=dtype, device=x.device)
    lev = torch.empty((B, T), dtype=torch.int64, device=x.device) if return_levels else None
    _lib.call("wnb200_siggen_onehot", ops._DT[dtype], B, T, num_levels, ops._p(x), ops._p(oh), ops._p(lev),
              ops._stream())
    return (oh, lev) if return_levels else oh
