"""Pins wavenet_speech_b200/utils/signal_gen.py (and through it csrc/siggen.cu) to the REFERENCE's own generator code.
TEST INFRASTRUCTURE: run in the authoring container, where /root/reference is mounted; writes
tests/golden/siggen_reference.npz.

What runs here is the reference's code, unmodified: `RawSignalGenerator.gaussian_model_fn` / `nts_to_kmer` /
`random_upsample` (utils/raw_signal_generator.py:91-118,189-203) and `quantize_fn` / `one_hot_fn`
(utils/gaussian_kmer_model.py:78-96, the same law as utils/pore_model.py:58-96).  raw_signal_generator.py imports
h5py at module level (:21) only to open a reference genome in __init__; the module is imported with an empty stand-in
for h5py and the object is built without __init__ (the genome is replaced by bases ~ U{1..4}, as
gaussian_kmer_model.py:281 does).  numpy's global legacy generator is seeded, the reference draws from it, and the
same draws are replayed here (gamma durations, standard normals) so that the fixture holds the draws AND the
reference's outputs for them."""
import os
import sys
import types

import numpy as np

REF = os.environ.get("WNB200_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "siggen_reference.npz")


def main():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, REF)
    from utils import raw_signal_generator as RSG
    from utils import gaussian_kmer_model as GKM
    tmpl = np.load(os.path.join(REF, "utils", "r9.4_450bps.5mer.template.npz"))
    gen = object.__new__(RSG.RawSignalGenerator)
    gen.kmer_means, gen.kmer_stdvs = tmpl["means"], tmpl["stdvs"]
    gen.duration_shape, gen.duration_rate, gen.sample_rate = 2.461964, 587.2858, 800.
    out = {}
    for case, (seed, nb) in enumerate([(7, 300), (8, 57), (9, 5)]):
        rs = np.random.RandomState(1000 + seed)
        bases = rs.randint(1, 5, size=nb)                              # gaussian_kmer_model.py:281 draws U{1..4}
        np.random.seed(seed)
        sig = gen.gaussian_model_fn(bases)                             # the reference, with numpy's global generator
        # replay the draws the reference just consumed
        np.random.seed(seed)
        n_k = nb - 4
        g = np.random.gamma(gen.duration_shape, np.reciprocal(gen.duration_rate), size=(n_k,))
        reps = (g * gen.sample_rate).astype(np.int32)
        reps = reps + (reps == 0).astype(np.int32)
        z = np.random.standard_normal(size=int(reps.sum()))
        # the reference's own k-mer ids, by its own window function
        from scipy.ndimage import generic_filter
        kmers = generic_filter(bases, gen.nts_to_kmer, size=(5,), mode="constant")[2:-2].astype(np.int64)
        seq = np.repeat(kmers, reps)
        replay = gen.kmer_means[seq].astype(np.float64) + gen.kmer_stdvs[seq].astype(np.float64) * z
        assert sig.shape == replay.shape and np.array_equal(sig, replay), "draw replay does not reproduce the reference"
        out["c%d/bases" % case], out["c%d/kmers" % case] = bases.astype(np.int64), kmers
        out["c%d/reps" % case], out["c%d/z" % case], out["c%d/sig" % case] = reps, z, sig
        if sig.shape[0] >= 8:
            q = object.__new__(GKM.GaussianModel) if hasattr(GKM, "GaussianModel") else None
            if q is None:                                               # class name differs across revisions: find it
                cls = [v for v in vars(GKM).values() if isinstance(v, type) and hasattr(v, "quantize_fn")][0]
                q = object.__new__(cls)
            q.num_levels = 256
            _mu = 256.0
            q.quant_law = np.vectorize(lambda x: (np.sign(x) * (np.log(1 + _mu * np.abs(x)) * np.reciprocal(np.log(1 + _mu)))))
            q.quant_levels = np.linspace(-1.0, 1.0, num=256)
            lev = q.quantize_fn(sig)
            out["c%d/levels" % case] = lev.astype(np.int64)
            out["c%d/onehot_argmax" % case] = q.one_hot_fn(lev).argmax(0).astype(np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
