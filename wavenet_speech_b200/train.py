"""The WaveNet-CTC training step (reference: legacy_code/train.py:24-61) on the device kernels.

Same arithmetic as the reference's `train_step`: cross-entropy of the WaveNet's next-sample prediction summed over
time with each frame averaged over the batch, CTC (blank = 0, labels shifted by +1) summed over the batch, the joint
`xe / T + ctc / T'` back-propagated, one optimiser step.  What differs is where it runs: the T-iteration Python loop
of CrossEntropyLoss calls is one fused log-softmax + NLL launch, the warp-ctc host hop is the device CTC, and with
`world > 1` (one process per GPU, batch sharded) the gradients are all-reduced before the optimiser step with the
batch-mean term pre-scaled by 1/world (sharding.joint_loss_for_backward).
"""
import torch

from . import functional as WF, ops, sharding


def train_step(wavenet, ctcnet, sig, seq, lengths, opt, batch_size=None, averaged=True, world=1, group=None,
               labels_are_zero_based=True):
    """sig: (B, levels, T) one-hot signal (bf16 for the tensor-core path); seq: flat integer labels of all reads
    (0-based as in the reference, shifted by +1 here so that 0 is the blank, train.py:44-45); lengths: labels per
    read.  Returns device scalars (xe, ctc, joint) -- averaged per frame like the reference unless averaged=False.
    No host synchronisation happens in here."""
    B = sig.shape[0] if batch_size is None else batch_size
    T = sig.shape[2]
    opt.zero_grad(set_to_none=True)
    pred = wavenet(sig[:, :, 0:-1])                                        # train.py:30
    trans = ctcnet(pred)                                                   # train.py:33
    # train.py:36 -- the per-frame argmax of the WHOLE signal, then the shift: the same integers as argmax(sig[:, :, 1:])
    # without copying the (B, levels, T-1) slice, and on rows that stay 16-byte aligned
    dense = ops.argmax_channels(sig)[:, 1:].contiguous()
    xe = WF.cross_entropy_sum(pred, dense) / B                             # train.py:37-39 (batch mean per frame)
    labels = seq.to(device=sig.device, dtype=torch.int32)
    if labels_are_zero_based:
        labels = labels + 1                                                # train.py:44-45: <0> == <BLANK>
    ctc = WF.ctc_loss_sum(trans, labels, lengths, layout="bct")            # train.py:42-46, read in place
    Tc = trans.shape[2]
    loss = sharding.joint_loss_for_backward(xe, ctc, T, Tc, world)         # train.py:50-53
    loss.backward()
    if world > 1:
        sharding.allreduce_gradients(list(wavenet.parameters()) + list(ctcnet.parameters()), group=group)
    opt.step()
    if averaged:
        return xe.detach() / T, ctc.detach() / Tc, xe.detach() / T + ctc.detach() / Tc
    return xe.detach(), ctc.detach(), xe.detach() + ctc.detach()
