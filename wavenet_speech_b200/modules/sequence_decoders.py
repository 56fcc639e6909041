"""Greedy sequence decoders (reference: modules/sequence_decoders.py:9-41, Decoder.py:20-35).  The beam-search
decoder of the reference is host-side Python and outside the hot path."""
import torch

from .. import ops

_LOOKUP = {0: '', 1: 'A', 2: 'G', 3: 'C', 4: 'T'}


def argmax_decode(logits):
    """logits (batch, sequence, logit) -> LongTensor (batch, sequence) of per-frame argmax labels; no repeat
    collapse (reference sequence_decoders.py:9-23).  The permuted view of a (batch, logit, sequence) network output
    is read in place."""
    return ops.frame_argmax(logits, layout="btc")


def labels2strings(labels, lookup=_LOOKUP):
    """(batch, sequence) integer labels -> list of strings through `lookup` (reference sequence_decoders.py:26-41)."""
    rows = labels.detach().cpu().tolist()
    return ["".join(lookup[int(ix)] for ix in row) for row in rows]


def greedy_ctc_decode(logits, lengths=None, blank=0, lookup=_LOOKUP):
    """logits (batch, num_labels, sequence) as the networks emit them -> list of decoded strings: argmax, collapse
    repeats, drop blanks (what ipynbs/Size 1 Pore Model Check.ipynb cell 24 does by hand), all on the device; only
    the packed labels cross to the host."""
    lab, n = ops.ctc_greedy_decode(logits, lengths, blank, layout="bct")
    lab, n = lab.cpu(), n.cpu().tolist()
    return ["".join(lookup[int(ix)] for ix in lab[b, :n[b]].tolist()) for b in range(lab.shape[0])]


class Decoder(object):
    """decoder='argmax' arm of the reference's Decoder wrapper (Decoder.py:7-35): decode(logits (batch, num_labels,
    sequence)) -> (None, list of strings), per-frame argmax with blanks mapped to ''."""

    def __init__(self, decoder='argmax', batch_size=None, num_labels=None, **_unused):
        if decoder != 'argmax':
            raise NotImplementedError("only the argmax decoder is on the hot path; beam search is host-side Python")
        self.decoder_type = decoder

    def decode(self, logits):
        return None, labels2strings(argmax_decode(logits.permute(0, 2, 1)))
