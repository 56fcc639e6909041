"""
CPU oracle for the wavenet-speech hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (state_dict in, tensor out) CPU restatement of the
reference's algorithm for the dilated residual stack, RawCTCNet and the
WaveNetClassifier.  It exists so that `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` have something to check
and time the CUDA path against.  Nothing under `wavenet_speech_b200/` may import
it; the product path has no CPU fallback.

Parity status: PINNED.  `oracle/gen_golden.py` (run in the authoring container,
where /root/reference is mounted) builds the reference's own nn.Modules, runs
them on seeded inputs and stores weights/inputs/outputs under `tests/golden/`;
`tests/test_oracle_golden.py` requires this file to reproduce every stored
output bit-for-bit in fp32.  The two CTC known-answer values the reference holds
(tests/test_classifier.py:53-59 -> 2.4628, ipynbs/CTC Overfit.ipynb cell 27 ->
1.4519) pin the CTC stand-in (`ctc_loss_sum`).

Every function cites the reference file:line it follows (paths relative to the
reference repository root).  Weights are addressed by the reference's own
state_dict keys so that a reference checkpoint can be fed in directly.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# modules/conv_ops.py
# ----------------------------------------------------------------------------
def autopad(k, d):
    """modules/conv_ops.py:104-116 -- ceil((k-1)*d/2)."""
    total = (k - 1) * d
    if total % 2 == 1:
        return (total - 1) // 2 + 1
    return total // 2


def tap_offsets(k, d, causal):
    """Time offset read by tap j: y[t] += W[:,:,j] @ x[t + off_j].

    Causal: conv1d(padding=(k-1)d) truncated to T  (conv_ops.py:28-34,43-44)
            => off_j = j*d - (k-1)*d.
    Non-causal: conv1d(padding=autopad(k,d)) truncated to T (conv_ops.py:62-68,78-79)
            => off_j = j*d - autopad(k,d).
    """
    pad = (k - 1) * d if causal else autopad(k, d)
    return [j * d - pad for j in range(k)]


def causal_conv1d(x, w, b, d):
    """modules/conv_ops.py:39-44 (CausalConv1d.forward)."""
    k = w.shape[2]
    y = F.conv1d(x, w, b, stride=1, padding=(k - 1) * d, dilation=d)
    return y[:, :, 0:x.shape[2]]


def noncausal_conv1d(x, w, b, d):
    """modules/conv_ops.py:74-79 (NonCausalConv1d.forward)."""
    k = w.shape[2]
    y = F.conv1d(x, w, b, stride=1, padding=autopad(k, d), dilation=d)
    return y[:, :, 0:x.shape[2]]


def conv1d_taps_numpy(x, w, b, offsets):
    """Independent (slow) restatement of a dilated conv as an explicit tap sum with
    zero padding at the true sequence ends.  Used by the tests to pin `tap_offsets`
    against F.conv1d; small cases only."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    B, _, T = x.shape
    M = w.shape[0]
    y = np.zeros((B, M, T), dtype=np.float64)
    for j, off in enumerate(offsets):
        for t in range(T):
            s = t + off
            if 0 <= s < T:
                y[:, :, t] += x[:, :, s] @ w[:, :, j].T
    if b is not None:
        y += np.asarray(b, dtype=np.float64)[None, :, None]
    return y


def reshape_in(seq):
    """modules/conv_ops.py:91-94."""
    N, C, L = seq.shape
    return seq.permute(0, 2, 1).contiguous().view(N * L, C), (N, L)


def reshape_out(seq, dims):
    """modules/conv_ops.py:97-101."""
    N, L = dims
    return seq.view(N, L, -1).permute(0, 2, 1).contiguous()


def channel_softmax(x):
    """modules/wavenet.py:108-109 (softmax over channels of an NCL tensor via
    reshape_in -> F.softmax on a 2-D tensor (implicit dim=1) -> reshape_out)."""
    flat, axes = reshape_in(x)
    return reshape_out(F.softmax(flat, dim=1), axes)


# ----------------------------------------------------------------------------
# modules/block.py
# ----------------------------------------------------------------------------
def gated_activation(a, b):
    """modules/block.py:184-185."""
    return torch.mul(torch.tanh(a), torch.sigmoid(b))


def bf16_storage(x):
    """Round to bf16 with a straight-through gradient.  NOT part of the reference: passed as `q` to the
    forward functions below it reproduces where a bf16 kernel pipeline STORES activations between kernels
    (gate, residual stream, head activations), so that gradient parity tests are not dominated by LeakyReLU /
    rounding decisions flipping between an fp32 and a bf16 evaluation.  q=None is the reference, exactly."""
    return x + (x.detach().bfloat16().float() - x.detach())


def _id(x):
    return x


def residual_block(sd, prefix, x, d, causal, q=None):
    """modules/block.py:54-82 (ResidualBlock.forward) -> (residual_out, skip_out)."""
    q = q or _id
    conv = causal_conv1d if causal else noncausal_conv1d
    a = conv(x, sd[prefix + "conv_tanh.conv1d.weight"], sd[prefix + "conv_tanh.conv1d.bias"], d)
    b = conv(x, sd[prefix + "conv_sigmoid.conv1d.weight"], sd[prefix + "conv_sigmoid.conv1d.bias"], d)
    act = q(gated_activation(a, b))
    res1x1 = F.conv1d(act, sd[prefix + "conv1x1_residual.weight"], sd[prefix + "conv1x1_residual.bias"])
    skip = F.conv1d(act, sd[prefix + "conv1x1_skip.weight"], sd[prefix + "conv1x1_skip.bias"])
    flat, axes = reshape_in(x)
    proj = reshape_out(F.linear(flat, sd[prefix + "residual_proj.weight"], sd[prefix + "residual_proj.bias"]), axes)
    return q(res1x1 + proj), skip


def _output_stack(sd, prefix, x, q=None):
    """modules/wavenet.py:67-71 / raw_ctcnet.py:84-88 / classifier.py:70-74:
    LeakyReLU(0.01) -> 1x1 -> LeakyReLU(0.01) -> 1x1."""
    q = q or _id
    h = q(F.leaky_relu(q(x), 0.01))
    h = F.conv1d(h, sd[prefix + "1.weight"], sd[prefix + "1.bias"])
    h = q(F.leaky_relu(h, 0.01))
    return F.conv1d(h, sd[prefix + "3.weight"], sd[prefix + "3.bias"])


# ----------------------------------------------------------------------------
# modules/wavenet.py
# ----------------------------------------------------------------------------
def wavenet_forward(sd, signal, layers, softmax=True, q=None):
    """modules/wavenet.py:88-111 (WaveNet.forward).  `layers` = [(c_in,c_out,k,d)]."""
    out = (q or _id)(causal_conv1d(signal, sd["entry_conv1d.conv1d.weight"], sd["entry_conv1d.conv1d.bias"], 1))
    out_dim = sd["bottlenecks.0.weight"].shape[0]
    skips = signal.new_zeros(signal.shape[0], out_dim, signal.shape[2])
    for l, (_ci, _co, _k, d) in enumerate(layers):
        out, skip = residual_block(sd, "convolutions.%d." % l, out, d, True, q)
        skips = skips + F.conv1d(skip, sd["bottlenecks.%d.weight" % l], sd["bottlenecks.%d.bias" % l])
    y = _output_stack(sd, "output_stack.", skips, q)
    if not softmax:
        return y
    return channel_softmax(y)


# ----------------------------------------------------------------------------
# modules/raw_ctcnet.py
# ----------------------------------------------------------------------------
def raw_ctcnet_forward(sd, seq, layers, input_dilation=1, positions=False, softmax=True, causal=False, q=None):
    """modules/raw_ctcnet.py:117-153 (RawCTCNet.forward).  Output length is
    T + feature_kwidth - 1: the featuriser pads by fk-1 and never truncates
    (raw_ctcnet.py:58)."""
    w0 = sd["feature_layer.0.weight"]
    fk = w0.shape[2]
    out = F.conv1d(seq, w0, sd["feature_layer.0.bias"], padding=fk - 1)
    q_ = q or _id
    out = q_(F.leaky_relu(out, 0.01))
    out = F.conv1d(out, sd["feature_layer.2.weight"], sd["feature_layer.2.bias"])
    out = q_(F.leaky_relu(out, 0.01))
    if positions:
        # raw_ctcnet.py:131-135
        incr = torch.arange(0., out.shape[2], dtype=out.dtype)
        pos = F.conv1d(incr.unsqueeze(0).unsqueeze(1), sd["positions_conv1x1.0.weight"], sd["positions_conv1x1.0.bias"])
        out = out + F.hardtanh(pos)
    out_dim = sd["input_skip_bottleneck.weight"].shape[0]
    skips = out.new_zeros(out.shape[0], out_dim, out.shape[2])
    out, skip = residual_block(sd, "input_block.", out, input_dilation, causal, q)
    skips = skips + F.conv1d(skip, sd["input_skip_bottleneck.weight"], sd["input_skip_bottleneck.bias"])
    for l, (_ci, _co, _k, d) in enumerate(layers):
        out, skip = residual_block(sd, "convolutions.%d." % l, out, d, causal, q)
        skips = skips + F.conv1d(skip, sd["bottlenecks.%d.weight" % l], sd["bottlenecks.%d.bias" % l])
    y = _output_stack(sd, "output_block.", skips, q)
    if not softmax:
        return y
    return channel_softmax(y)


# ----------------------------------------------------------------------------
# modules/classifier.py
# ----------------------------------------------------------------------------
def classifier_forward(sd, seq, layers, pool_kernel_size=2, input_dilation=1, softmax=True, q=None):
    """modules/classifier.py:91-120 (WaveNetClassifier.forward)."""
    out = (q or _id)(F.avg_pool1d(seq, kernel_size=pool_kernel_size, padding=0))   # classifier.py:53,102
    out_dim = sd["input_skip_bottleneck.weight"].shape[0]
    skips = out.new_zeros(out.shape[0], out_dim, out.shape[2])
    out, skip = residual_block(sd, "input_block.", out, input_dilation, False, q)
    skips = skips + F.conv1d(skip, sd["input_skip_bottleneck.weight"], sd["input_skip_bottleneck.bias"])
    for l, (_ci, _co, _k, d) in enumerate(layers):
        out, skip = residual_block(sd, "convolutions.%d." % l, out, d, False, q)
        skips = skips + F.conv1d(skip, sd["bottlenecks.%d.weight" % l], sd["bottlenecks.%d.bias" % l])
    y = _output_stack(sd, "output_block.", skips, q)
    if not softmax:
        return y
    return channel_softmax(y)


# ----------------------------------------------------------------------------
# modules/layernorm.py, modules/linear_conv_ops.py
# ----------------------------------------------------------------------------
def layernorm(x, gamma, beta, dim=1, eps=1e-6):
    """modules/layernorm.py:25-28: unbiased std, eps added to the std."""
    mean = x.mean(dim, keepdim=True).expand_as(x)
    std = x.std(dim, keepdim=True).expand_as(x)
    return gamma.expand_as(x) * (x - mean) / (std + eps) + beta.expand_as(x)


def get_ker_ixs(d, k):
    """modules/linear_conv_ops.py:112-123."""
    total = k * d - (d - 1)
    return [i for i in range(total) if i % d == 0]


def linear_conv1d_linear(frame, w, b, d, keep_dims=False):
    """modules/linear_conv_ops.py:39-68 (LinearConv1d.linear): one output frame."""
    k = w.shape[2]
    rf = k + (d - 1) * (k - 1)
    assert frame.shape[2] == rf
    ixs = get_ker_ixs(d, k)
    out = F.linear(frame[:, :, ixs].reshape(frame.shape[0], frame.shape[1] * k),
                   w.reshape(w.shape[0], w.shape[1] * w.shape[2]), b)
    return out.unsqueeze(2) if keep_dims else out


def multiplicative_unit(sd, prefix, h, d):
    """modules/block.py:213-220 (MultiplicativeUnit.forward)."""
    g1 = torch.sigmoid(causal_conv1d(h, sd[prefix + "gate1.conv1d.weight"], sd[prefix + "gate1.conv1d.bias"], d))
    g2 = torch.sigmoid(causal_conv1d(h, sd[prefix + "gate2.conv1d.weight"], sd[prefix + "gate2.conv1d.bias"], d))
    g3 = torch.sigmoid(causal_conv1d(h, sd[prefix + "gate3.conv1d.weight"], sd[prefix + "gate3.conv1d.bias"], d))
    u = torch.tanh(causal_conv1d(h, sd[prefix + "update.conv1d.weight"], sd[prefix + "update.conv1d.bias"], d))
    return g1.mul(torch.tanh(g2.mul(h) + g3.mul(u)))


def residual_relu_block(sd, prefix, x, d):
    """modules/block.py:130-173 (ResidualReLUBlock.forward): x + [LN, ReLU, 1x1 (C -> C/2), LN, ReLU, causal conv
    (k, d), LN, ReLU, 1x1 (C/2 -> C)](x).  `stack.{i}` are the reference's nn.Sequential indices."""
    s = prefix + "stack."
    h = torch.relu(layernorm(x, sd[s + "0.gamma"], sd[s + "0.beta"]))
    h = F.conv1d(h, sd[s + "2.weight"], sd[s + "2.bias"])
    h = torch.relu(layernorm(h, sd[s + "3.gamma"], sd[s + "3.beta"]))
    h = causal_conv1d(h, sd[s + "5.conv1d.weight"], sd[s + "5.conv1d.bias"], d)
    h = torch.relu(layernorm(h, sd[s + "6.gamma"], sd[s + "6.beta"]))
    h = F.conv1d(h, sd[s + "8.weight"], sd[s + "8.bias"])
    return x + h


def residual_mu_block(sd, prefix, x, d):
    """modules/block.py:86-126 (ResidualMUBlock.forward): x + [LN, ReLU, 1x1 (C -> C/2), LN, ReLU, MU(k, d), MU(1),
    1x1 (C/2 -> C)](x)."""
    s = prefix + "stack."
    h = torch.relu(layernorm(x, sd[s + "0.gamma"], sd[s + "0.beta"]))
    h = F.conv1d(h, sd[s + "2.weight"], sd[s + "2.bias"])
    h = torch.relu(layernorm(h, sd[s + "3.gamma"], sd[s + "3.beta"]))
    h = multiplicative_unit(sd, s + "5.", h, d)
    h = multiplicative_unit(sd, s + "6.", h, 1)
    h = F.conv1d(h, sd[s + "7.weight"], sd[s + "7.bias"])
    return x + h


def linear_conv1d_stream(seq, w, b, d):
    """LinearConv1d.linear (modules/linear_conv_ops.py:39-68) on every causal window of `seq` (zeros before the
    first frame): the frame-at-a-time evaluation a decoder asks for.  -> (batch, c_out, T)."""
    k = w.shape[2]
    rf = k + (d - 1) * (k - 1)
    padded = torch.cat([seq.new_zeros(seq.shape[0], seq.shape[1], rf - 1), seq], 2)
    return torch.stack([linear_conv1d_linear(padded[:, :, t:t + rf], w, b, d) for t in range(seq.shape[2])], 2)


# ----------------------------------------------------------------------------
# losses / decode (legacy_code/train.py, modules/sequence_decoders.py)
# ----------------------------------------------------------------------------
def xe_loss_sum_over_time(pred, dense_target):
    """legacy_code/train.py:36-39: sum over t of CrossEntropyLoss()(pred[:,:,t], tgt[:,t]);
    each term is a mean over the batch.  Vectorised (identical value up to fp32
    summation order)."""
    B = pred.shape[0]
    return F.cross_entropy(pred, dense_target, reduction="sum") / B


def ctc_loss_sum(activations_tbc, flat_labels, act_lengths, label_lengths, dtype=torch.float32):
    """Stand-in for warpctc_pytorch.CTCLoss as the reference calls it
    (legacy_code/train.py:42-46): pre-softmax activations (T,B,C), flattened int
    labels with 0 = blank, summed over the batch.  warp-ctc is an un-vendored,
    un-pinned third-party dependency (README.md:15-16); its published semantics
    are softmax-inside + sum reduction, reproduced here with torch's CTC."""
    logp = F.log_softmax(activations_tbc.to(dtype), dim=2)     # dtype=float64: the precision yardstick of the tests
    return F.ctc_loss(logp, flat_labels, act_lengths, label_lengths, blank=0, reduction="sum",
                      zero_infinity=False)


def train_step_losses(wsd, csd, sig, seq_flat, lengths, wave_layers, cls_layers, pool):
    """legacy_code/train.py:24-55 up to (not including) backward/opt.step: returns
    (avg_xe, avg_ctc, avg_joint, wavenet_pred, transcription).  `seq_flat` holds
    labels already shifted so that 0 = blank (train.py:44-45 adds one)."""
    pred = wavenet_forward(wsd, sig[:, :, 0:-1], wave_layers, softmax=False)
    trans = classifier_forward(csd, pred, cls_layers, pool_kernel_size=pool, softmax=False)
    dense = torch.max(sig[:, :, 1:], dim=1)[1]
    xe = xe_loss_sum_over_time(pred, dense)
    probs = trans.permute(2, 0, 1).contiguous()
    B = sig.shape[0]
    prob_lengths = torch.full((B,), probs.shape[0], dtype=torch.int32)
    ctc = ctc_loss_sum(probs, seq_flat, prob_lengths, lengths)
    avg_xe = xe / sig.shape[2]
    avg_ctc = ctc / trans.shape[2]
    return avg_xe, avg_ctc, avg_xe + avg_ctc, pred, trans


def argmax_decode(logits_bsl):
    """modules/sequence_decoders.py:9-23: per-frame argmax over the last dim of a
    (batch, sequence, logit) tensor; no repeat collapse."""
    return torch.max(logits_bsl, dim=2)[1]


def collapse_decode(labels_1d, blank=0):
    """Collapse repeats then drop blanks (what ipynbs/Size 1 Pore Model Check.ipynb
    cell 24 does by hand on top of argmax_decode)."""
    out = []
    prev = None
    for v in labels_1d.tolist():
        if v != prev and v != blank:
            out.append(v)
        prev = v
    return out


# ----------------------------------------------------------------------------
# algorithmic work
# ----------------------------------------------------------------------------
def block_flops(cin, cout, k, out_dim):
    """FLOP per timestep of ResidualBlock + bottleneck as written (MAC = 2)."""
    return 2 * (2 * k * cin * cout + 2 * cout * cout + cin * cout + cout * out_dim)


def wavenet_flops_per_timestep(in_dim, entry_kwidth, layers, out_dim):
    f = 2 * entry_kwidth * in_dim * layers[0][0]
    for (ci, co, k, _d) in layers:
        f += block_flops(ci, co, k, out_dim)
    f += 2 * (out_dim * out_dim) * 2
    return f
