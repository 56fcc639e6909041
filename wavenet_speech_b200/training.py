"""Tensor-core training path: forward that keeps what backward needs, and the backward itself, for the
residual stack and the networks built on it (WaveNet, WaveNetClassifier) -- the WaveNet-CTC train step of
reference legacy_code/train.py:24-61 (BASELINE config 3).

Everything is NLC bf16 between kernels, fp32 accumulation in TMEM, fp32 parameter gradients.  Per block
(reference block.py:54-82 + the bottleneck of wavenet.py:100), with row vectors per frame:

  forward   gate = th * sg,  th = tanh(Wt (*) x + bt),  sg = sigmoid(Ws (*) x + bs)      fused block kernel,
            res  = Wres gate + Wproj x + b ;  skips += Wbn (Wskip gate + bskip) + bbn     saves gate, th, sg
  backward  dgate = Wres^T dres + (Wbn Wskip)^T dskips                 two-source dense contraction
            dab   = [dgate sg (1-th^2) ; dgate th sg (1-sg)]           gate-backward kernel
            dx    = sum_j [Wt_j ; Ws_j]^T dab(t - off_j) + Wproj^T dres   two-source dense, negated tap offsets
            dWt_j, dWs_j = dab (x) x(t + off_j);  dWres = dres (x) gate;  dWproj = dres (x) x   time-contraction
            M = dskips (x) gate;  dWskip = Wbn^T M;  dWbn = M Wskip^T + colsum(dskips) bskip^T   (weight space)

`dskips` (gradient of the running skip sum) is the same tensor for every layer.
"""
import ctypes

import torch

from . import _lib, fastpath as FP, ops


def _t(w):
    return w.detach().float().t()


class _ZeroPool(object):
    """fp32 zero-initialised scratch handed out in slices: ONE memset per backward pass instead of one fill launch per
    accumulator (weight-gradient tiles, bias sums) -- several hundred 2-microsecond launches per step otherwise."""
    current = None

    def __init__(self, dev, nfloats):
        self.buf = torch.zeros(int(nfloats), dtype=torch.float32, device=dev)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_al = (n + 63) // 64 * 64                      # 256-byte aligned slices (TMA reduce-add targets)
        if self.off + n_al > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float32, device=self.buf.device)
        out = self.buf[self.off:self.off + n].view(*shape)
        self.off += n_al
        return out


def _zeros(n, dev):
    pool = _ZeroPool.current
    if pool is not None and pool.buf.device == torch.device(dev):
        return pool.take(n)
    return torch.zeros(n, dtype=torch.float32, device=dev)


class zero_pool(object):
    """with zero_pool(device, nfloats): ... -> `_zeros` and the weight-gradient launches draw from one buffer."""

    def __init__(self, dev, nfloats):
        self.dev, self.n = dev, nfloats

    def __enter__(self):
        self.prev = _ZeroPool.current
        _ZeroPool.current = _ZeroPool(self.dev, self.n)

    def __exit__(self, *exc):
        _ZeroPool.current = self.prev


# True: the d(gate) contraction applies the gate's backward in its epilogue (wnb200_dense_t.gb_gate / gb_sg) instead of a
# separate elementwise launch.  Measured at the config-3 layer shape (scripts/bench_gatebwd.py, 32 x 16383 x 256): fused
# 0.39 ms, DRAM traffic at the algorithmic minimum (1.07 GB read + 0.50 GB written) -- but the contraction + the
# elementwise launch take 0.145 + 0.207 = 0.35 ms: the elementwise kernel runs 64 warps per SM at 6.4 TB/s, the fused
# epilogue has the kernel's 8 epilogue warps per SM to do the same arithmetic and is latency-bound there (0.2 IPC).
# So the two-launch form stays the default; the fused form is kept, tested, for shapes where launches dominate.
FUSE_GATE_BWD = False


def colsum(x_nlc, out=None):
    """fp32 column sums of an NLC bf16 tensor [B, T, C] (accumulates into `out`)."""
    B, T, C = x_nlc.shape
    if out is None:
        out = _zeros(C, x_nlc.device)
    _lib.call("wnb200_colsum_nlc", B * T, C, ops._p(x_nlc), ops._p(out), ops._stream())
    return out


def gate_bwd_nlc(dgate, th, sg, want_bias=False, th_is_gate=False, out=None, dbias=None):
    """dab = [dgate sg (1-th^2) ; dgate th sg (1-sg)]; with want_bias also its fp32 column sums [2C].
    th_is_gate: `th` holds the gate tanh * sigmoid (what the training forward keeps): tanh = gate / sigmoid.
    out / dbias: write into / accumulate into caller-owned tensors (a batch walked in chunks)."""
    B, T, C = dgate.shape
    dab = torch.empty((B, T, 2 * C), dtype=torch.bfloat16, device=dgate.device) if out is None else out
    if want_bias and dbias is None:
        dbias = _zeros(2 * C, dgate.device)
    _lib.call("wnb200_gate_bwd_nlc_from_gate" if th_is_gate else "wnb200_gate_bwd_nlc", B * T, C, ops._p(dgate),
              ops._p(th), ops._p(sg), ops._p(dab), ops._p(dbias), ops._stream())
    return (dab, dbias) if want_bias else dab


def leaky_bwd(dy, ref):
    """dy * (ref > 0 ? 1 : 0.01), any layout (ref = the LeakyReLU's input or output: same sign)."""
    return ops.leaky_bwd(dy, ref)


def wgrad_multi(g, rows, srcs):
    """For every (x, off) in srcs: sum_{b,t} g[b,t,m] x[b,t+off,n], m < rows -> list of fp32 [rows, N].
    256 rows of g and two x tensors per launch (the g tiles are read once for both)."""
    N = srcs[0][0].shape[2]
    outs = [[] for _ in srcs]
    for m0 in range(0, rows, 256):
        nr = min(256, rows - m0)
        for i in range(0, len(srcs), 2):
            pair = srcs[i:i + 2]
            dw = FP.wgrad2(g, [q[0] for q in pair], [q[1] for q in pair], m0, dw=_zeros(256 * len(pair) * N, g.device).view(256, len(pair) * N))
            for q in range(len(pair)):
                outs[i + q].append(dw[:nr, q * N:(q + 1) * N])
    return [o[0] if len(o) == 1 else torch.cat(o, 0) for o in outs]


def wgrad_group(reqs):
    """reqs: list of (g, rows, srcs, dw_into) with wgrad_multi's meaning, all over the same frames: every 256-row x
    (one or two sources) product of every request goes into as few launches as possible (FP.wgrad_jobs), so that the
    operands the requests share -- x, the gate, the gradients -- are read from HBM once.  dw_into (optional, one source,
    rows == 256): a zeroed fp32 [256, N] tensor that receives the product in place.  Returns wgrad_multi's lists."""
    jobs, outs = [], []
    for g, rows, srcs, dw_into in reqs:
        N = srcs[0][0].shape[2]
        o = [[] for _ in srcs]
        for m0 in range(0, rows, 256):
            nr = min(256, rows - m0)
            for i in range(0, len(srcs), 2):
                pair = srcs[i:i + 2]
                if dw_into is not None:
                    assert len(srcs) == 1 and rows == 256
                    dw = dw_into
                else:
                    dw = _zeros(256 * len(pair) * N, g.device).view(256, len(pair) * N)
                jobs.append((g, m0, [q[0] for q in pair], [q[1] for q in pair], dw))
                for q in range(len(pair)):
                    o[i + q].append(dw[:nr, q * N:(q + 1) * N])
        outs.append(o)
    FP.wgrad_jobs(jobs)
    return [[p[0] if len(p) == 1 else torch.cat(p, 0) for p in o] for o in outs]


def wgrad_rows(g, x, off, rows, N):
    return wgrad_multi(g, rows, [(x, off)])[0]


class Stack(object):
    """The blocks + bottlenecks of one network with their forward / backward weight packs: one `wnb200_pack_block`
    launch per layer (pipelined row order + the transposed matrices of the data gradients; fold product on the device)."""

    def __init__(self, blocks, bottlenecks):
        self.blocks, self.necks = list(blocks), list(bottlenecks)
        self.fwd = [FP.pack_block(b, n, precise=False, bwd=True, natural=False) for b, n in zip(self.blocks, self.necks)]
        self.bwd = self.fwd                   # same dicts: wdg / wdg_skip / wdx / wdx_taps / k / C live next to w1h ...
        self._skip = None

    def skip_pack(self):
        if self._skip is None:
            self._skip = FP.skip_pack(self.fwd)
        return self._skip

    def params(self):
        """Per layer, in this order: wt, bt, ws, bs, wres, bres, wskip, bskip, wproj, bproj, wbn, bbn."""
        out = []
        for b, n in zip(self.blocks, self.necks):
            out += [b.conv_tanh.conv1d.weight, b.conv_tanh.conv1d.bias, b.conv_sigmoid.conv1d.weight,
                    b.conv_sigmoid.conv1d.bias, b.conv1x1_residual.weight, b.conv1x1_residual.bias,
                    b.conv1x1_skip.weight, b.conv1x1_skip.bias, b.residual_proj.weight, b.residual_proj.bias,
                    n.weight, n.bias]
        return out


def _pool_floats(stack):
    """Upper bound of the fp32 accumulators one backward pass draws from the zero pool."""
    C = stack.fwd[0]["C"]
    per_layer = 4 * 256 * 2 * C + 8 * 2 * C + 1024
    return (len(stack.fwd) + 4) * per_layer + 16 * 256 * 2 * 256


def stack_forward(h0, stack, skips):
    """Residual stack keeping (x, gate, sigmoid) of every layer -- backward recovers tanh as gate / sigmoid, so the third
    tensor round 1 wrote and read back per layer (2 x 268 MB at the config-3 shape) is gone.  Returns (saved list,
    skips_act): the last launch emits LeakyReLU(skip sum) as bf16 itself (the fp32 sum is not needed after the forward)."""
    saved = []
    h = h0
    n = len(stack.fwd)
    if FP.DEFER_SKIP and h0.shape[2] in (128, 256) and FP.RESBLOCK_VARIANT != 1:
        # deferred skip (resblock3_kernel): every layer stores its gate into a stack -- the activation backward keeps
        # anyway -- plus the sigmoid factor; the skip sum is ONE contraction over K = layers x channels afterwards, so
        # the fp32 running sum (1.07 GB of read-modify-write per layer at the config-3 shape) never exists in HBM.
        gates = torch.empty((n,) + tuple(h0.shape), dtype=h0.dtype, device=h0.device)
        for l, pk in enumerate(stack.fwd):
            last = l == n - 1
            sg = torch.empty_like(h)
            res = None if last else torch.empty_like(h)
            FP.resblock(h, pk, res, None, False, save=(None, None, sg), gate_out=gates[l])
            saved.append((h, gates[l], sg))
            h = res
        wcat, bsum = stack.skip_pack()
        return saved, FP.dense(gates, [0], wcat, bsum, h0.shape[2], leaky=1, nlayers=n)
    skips_act = torch.empty_like(h0) if FP.FUSE_FINAL else None
    for l, pk in enumerate(stack.fwd):
        last = l == n - 1
        act, sg = torch.empty_like(h), torch.empty_like(h)
        res = None if last else torch.empty_like(h)
        FP.resblock(h, pk, res, skips, l == 0, save=(act, None, sg), skips_act=skips_act if last else None)
        saved.append((h, act, sg))
        h = res
    return saved, skips_act


def stack_backward(stack, saved, dskips, need_dx0):
    """-> (dh0 or None, list of per-layer parameter gradients in Stack.params() order (fp32), fp32 column sums of dh0
    or None -- the bias gradient of whatever produced the stack's input, summed by the launch that wrote dh0)."""
    B, T, C = dskips.shape
    dev = dskips.device
    zb = _zeros(C, dev)
    csk = colsum(dskips)                                    # d(bottleneck bias), identical for every layer
    L = len(saved)
    csk_rows = csk.unsqueeze(0).repeat(L, 1)                # ... but every layer's parameter gets its own memory
    grads = [None] * L
    M_all = _zeros(L * C * C, dev).view(L, C, C)            # M[l] = dskips (x) gate_l, the fold's weight-space gradient
    dres = dres_cs = None
    for l in range(L - 1, -1, -1):
        x, act, sg = saved[l]
        pb, offs = stack.bwd[l], stack.fwd[l]["offsets"]
        k = pb["k"]
        dbab = _zeros(2 * C, dev)
        if not FUSE_GATE_BWD:
            dg = (FP.dense(dskips, [0], pb["wdg_skip"], zb, C) if dres is None
                  else FP.dense(dres, [0], pb["wdg"], zb, C, x2=dskips, offsets2=[0]))
            dab, dbab = gate_bwd_nlc(dg, act, sg, want_bias=True, th_is_gate=True)
            del dg
        elif dres is None:
            dab = FP.dense(dskips, [0], pb["wdg_skip"], zb, C, gate_bwd=(act, sg), colsum=dbab)
        else:
            dab = FP.dense(dres, [0], pb["wdg"], zb, C, x2=dskips, offsets2=[0], gate_bwd=(act, sg), colsum=dbab)
        dx = dx_cs = None
        if l > 0 or need_dx0:
            neg = [-o for o in offs]
            dx_cs = _zeros(C, dev)               # column sums of dx come out of the same launch (bias gradients below)
            if dres is None:
                dx = FP.dense(dab, neg, pb["wdx_taps"], zb, C, colsum=dx_cs)
            else:
                dx = FP.dense(dab, neg, pb["wdx"], zb, C, x2=dres, offsets2=[0], colsum=dx_cs)
        # all weight gradients of the block in one launch: k x [2C, C] for the taps, [gate | x] for the two 1x1s on the
        # residual path, and M[l] (in place when it is one 256-row tile)
        reqs = [(dab, 2 * C, [(x, offs[j]) for j in range(k)], None)]
        if dres is not None:
            reqs.append((dres, C, [(act, 0), (x, 0)], None))
        reqs.append((dskips, C, [(act, 0)], M_all[l] if C == 256 else None))
        got = wgrad_group(reqs)
        dwab = got[0]
        dwt = torch.stack([d[:C] for d in dwab], 2)
        dws = torch.stack([d[C:] for d in dwab], 2)
        dwres = dwproj = dbres = None
        if dres is not None:
            dwres, dwproj = got[1]
            dwres = dwres.unsqueeze(2)
            dbres = dres_cs
        if C != 256:
            M_all[l].copy_(got[-1][0])
        grads[l] = [dwt, dbab[:C], dws, dbab[C:], dwres, dbres, None, None, dwproj, dbres, None, csk_rows[l]]
        saved[l] = None                                     # free this layer's activations
        dres, dres_cs = dx, dx_cs
    # weight-space algebra of the folded skip -> bottleneck product: dWskip = Wbn^T M, dWbn = M Wskip^T + csk (x) bskip,
    # dbskip = Wbn^T csk -- every layer in ONE launch of wnb200_fold_grads (fp32; these were cuBLAS bmm calls in round 1)
    wbn_p = [n.weight for n in stack.necks]
    wsk_p = [b.conv1x1_skip.weight for b in stack.blocks]
    bsk_p = [b.conv1x1_skip.bias for b in stack.blocks]
    srcs, wdt = FP._sources(wbn_p + wsk_p + bsk_p)
    dwskip = torch.empty((L, C, C), dtype=torch.float32, device=dev)
    dwbn = torch.empty((L, C, C), dtype=torch.float32, device=dev)
    dbskip = torch.empty((L, C), dtype=torch.float32, device=dev)
    for l0 in range(0, L, 64):
        n = min(64, L - l0)
        sl = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts[l0:l0 + n]])
        _lib.call("wnb200_fold_grads", wdt, n, C, sl(srcs[0:L]), sl(srcs[L:2 * L]), sl(srcs[2 * L:3 * L]),
                  ops._p(M_all[l0:]), ops._p(csk), ops._p(dwskip[l0:]), ops._p(dwbn[l0:]), ops._p(dbskip[l0:]),
                  ops._stream())
    for l in range(L):
        grads[l][6], grads[l][7], grads[l][10] = dwskip[l].unsqueeze(2), dbskip[l], dwbn[l].unsqueeze(2)
    return dres, grads, dres_cs


def head_forward(skips, hd, out_dtype, softmax, skips_act=None):
    """LeakyReLU -> 1x1 -> LeakyReLU -> 1x1 [-> softmax]; returns (out NCL, skips_act, h1)."""
    B, T, C = skips.shape
    if skips_act is None:
        skips_act = FP.leaky_to_bf16(skips)
    h1 = FP.dense(skips_act, [0], hd["w1"], hd["b1"], C, leaky=1)
    out = torch.empty((B, hd["n_out"], T), dtype=out_dtype, device=skips.device)
    FP.dense(h1, [0], hd["w2"], hd["b2"], hd["n2"], mode=1, out=out, n_out=hd["n_out"], softmax=softmax)
    return out, skips_act, h1


def head_backward(dout, out, softmax, hb, skips_act, h1):
    """dout NCL [B, n_out, T] -> (dskips NLC bf16, [dw1, db1, dw3, db3])."""
    B, n_out, T = dout.shape
    C = h1.shape[2]
    dev = dout.device
    if softmax:
        dout = ops.softmax_bwd(out, dout.to(out.dtype))
    if hb["npad"] != n_out:
        pad = torch.zeros((B, hb["npad"], T), dtype=dout.dtype, device=dev)
        pad[:, :n_out] = dout
        dout = pad
    dl = FP.ncl_to_nlc_bf16(dout)                                                   # [B, T, npad]
    zb = _zeros(C, dev)
    dh1 = leaky_bwd(FP.dense(dl, [0], hb["w3t"], zb, C), h1)
    dw3 = wgrad_rows(dl, h1, 0, n_out, C).unsqueeze(2)
    db3 = colsum(dl)[:n_out]
    dskips = leaky_bwd(FP.dense(dh1, [0], hb["w1t"], zb, C), skips_act)
    dw1 = wgrad_rows(dh1, skips_act, 0, C, C).unsqueeze(2)
    db1 = colsum(dh1)
    return dskips, [dw1, db1, dw3, db3]


def _head_params(head):
    return [head[1].weight, head[1].bias, head[3].weight, head[3].bias]


def _cast(grads, params):
    """Gradients in the parameters' shapes / dtypes.  Every parameter gets its OWN memory: the same tensor handed to two
    parameters (one bias sum serves conv1x1_residual.bias and residual_proj.bias; the bottleneck biases of all layers
    share d(skip sum)) would be stolen as-is by AccumulateGrad, and clip_grad_norm_ / zero_grad(set_to_none=False) /
    gradient accumulation would then act on the shared memory once per alias."""
    seen, out = set(), []
    for g, p in zip(grads, params):
        if g is None:
            out.append(None)
            continue
        g = g.reshape(p.shape).to(p.dtype)
        if g.data_ptr() in seen:
            g = g.clone()
        seen.add(g.data_ptr())
        out.append(g)
    return tuple(out)


# --------------------------------------------------------------------------- WaveNet
def _wavenet_pack(model):
    C = model.layers[0][0]
    ec = model.entry_conv1d.conv1d
    w = ec.weight.detach().float()                                                  # [C, in_dim, k]
    return {"entry_w": FP._bf16(FP._taps_matrix(ec.weight)), "entry_b": ec.bias.detach().float().contiguous(),
            "entry_wt": FP._bf16(torch.cat([w[:, :, j].t() for j in range(w.shape[2])], 1)),   # [in_dim, k*C]
            "stack": Stack(model.convolutions, model.bottlenecks),
            "head": FP.pack_head(model.output_stack, C, bwd=True)}


def _wavenet_params(model, pk):
    ec = model.entry_conv1d.conv1d
    return [ec.weight, ec.bias] + pk["stack"].params() + _head_params(model.output_stack)


class _WaveNetTrain(torch.autograd.Function):
    """WaveNet.forward (reference wavenet.py:88-111) with its backward, both on the tensor-core kernels."""

    @staticmethod
    def forward(ctx, model, pk, signal, *params):
        C = model.layers[0][0]
        B, _, T = signal.shape
        x = FP.ncl_to_nlc_bf16(signal)
        offs = list(model.entry_conv1d.offsets)
        h0 = FP.dense(x, offs, pk["entry_w"], pk["entry_b"], C)
        skips = torch.empty((B, T, C), dtype=torch.float32, device=signal.device)
        saved, sk_act = stack_forward(h0, pk["stack"], skips)
        out, skips_act, h1 = head_forward(skips, pk["head"], signal.dtype, model.softmax, sk_act)
        del skips
        ctx.model, ctx.pk, ctx.params = model, pk, params
        ctx.keep = (x, offs, saved, skips_act, h1, out if model.softmax else None)
        ctx.in_dtype = signal.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        with zero_pool(dout.device, _pool_floats(ctx.pk["stack"])):
            return _WaveNetTrain._backward(ctx, dout)

    @staticmethod
    def _backward(ctx, dout):
        model, pk, params = ctx.model, ctx.pk, ctx.params
        x, offs, saved, skips_act, h1, out = ctx.keep
        ctx.keep = None
        dskips, ghead = head_backward(dout.contiguous(), out, model.softmax, pk["head"], skips_act, h1)
        dh0, gl, dbe = stack_backward(pk["stack"], saved, dskips, True)
        C, in_dim = dh0.shape[2], x.shape[2]
        dwe = torch.stack(wgrad_multi(dh0, C, [(x, o) for o in offs]), 2)
        dsignal = None
        if ctx.needs_input_grad[2]:
            dxn = FP.dense(dh0, [-o for o in offs], pk["entry_wt"], _zeros(in_dim, dh0.device), in_dim)
            dsignal = FP.nlc_to_ncl(dxn, ctx.in_dtype)
        grads = [dwe, dbe] + [g for layer in gl for g in layer] + ghead
        return (None, None, dsignal) + _cast(grads, params)


def wavenet_train_eligible(model, signal):
    C = model.layers[0][0]
    return (FP.tc_dtype_ok(signal) and signal.is_cuda and signal.dim() == 3 and C in (128, 256)
            and model.in_dim in (128, 256) and model.out_dim == C and FP._stack_ok(C, model.layers)
            and model.entry_kwidth <= 3 and signal.shape[0] > 0 and signal.shape[2] > 0)


def wavenet_forward_train(model, signal):
    ops.check_device()
    pk = FP._cached(model, "wavenet_train", lambda: _wavenet_pack(model))
    return _WaveNetTrain.apply(model, pk, signal, *_wavenet_params(model, pk))     # (time slices are read in place)


# --------------------------------------------------------------------------- WaveNetClassifier
def _classifier_pack(model):
    C = model.layers[0][0]
    return {"stack": Stack([model.input_block] + list(model.convolutions),
                           [model.input_skip_bottleneck] + list(model.bottlenecks)),
            "head": FP.pack_head(model.output_block, C, bwd=True)}


class _ClassifierTrain(torch.autograd.Function):
    """WaveNetClassifier.forward (reference classifier.py:91-120) with its backward on the tensor-core kernels."""

    @staticmethod
    def forward(ctx, model, pk, seq, *params):
        C = model.layers[0][0]
        pool = model.pool_kernel_size
        B, _, T = seq.shape
        To = T // pool
        h0 = torch.empty((B, To, C), dtype=torch.bfloat16, device=seq.device)
        _lib.call("wnb200_avgpool_ncl_to_nlc", ops._dt(seq), B, C, T, pool, ops._p(seq), _lib.ACT_BF16, ops._p(h0), ops._stream())
        skips = torch.empty((B, To, C), dtype=torch.float32, device=seq.device)
        saved, sk_act = stack_forward(h0, pk["stack"], skips)
        out, skips_act, h1 = head_forward(skips, pk["head"], seq.dtype, model.softmax, sk_act)
        ctx.model, ctx.pk, ctx.params = model, pk, params
        ctx.keep = (saved, skips_act, h1, out if model.softmax else None)
        ctx.T, ctx.in_dtype = T, seq.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        with zero_pool(dout.device, _pool_floats(ctx.pk["stack"])):
            return _ClassifierTrain._backward(ctx, dout)

    @staticmethod
    def _backward(ctx, dout):
        model, pk, params = ctx.model, ctx.pk, ctx.params
        saved, skips_act, h1, out = ctx.keep
        ctx.keep = None
        need_dx = ctx.needs_input_grad[2]
        dskips, ghead = head_backward(dout.contiguous(), out, model.softmax, pk["head"], skips_act, h1)
        dh0, gl, _ = stack_backward(pk["stack"], saved, dskips, need_dx)
        dseq = None
        if need_dx:
            B, _To, C = dh0.shape                 # layout change + AvgPool1d backward in one pass
            dseq = torch.empty((B, C, ctx.T), dtype=ctx.in_dtype, device=dh0.device)
            _lib.call("wnb200_avgpool_bwd_nlc_to_ncl", ops._DT[ctx.in_dtype], B, C, ctx.T, model.pool_kernel_size,
                      ops._p(dh0.contiguous()), ops._p(dseq), ops._stream())
        grads = [g for layer in gl for g in layer] + ghead
        return (None, None, dseq) + _cast(grads, params)


def classifier_train_eligible(model, seq):
    C = model.layers[0][0]
    return (FP.tc_dtype_ok(seq) and seq.is_cuda and seq.dim() == 3 and C in (128, 256) and model.in_dim == C
            and model.out_dim == C and FP._stack_ok(C, model.layers) and model.input_kernel_size <= 3
            and seq.shape[0] > 0 and seq.shape[2] // model.pool_kernel_size > 0)


def classifier_forward_train(model, seq):
    ops.check_device()
    pk = FP._cached(model, "classifier_train", lambda: _classifier_pack(model))
    params = pk["stack"].params() + _head_params(model.output_block)
    return _ClassifierTrain.apply(model, pk, seq.contiguous(), *params)


# --------------------------------------------------------------------------- RawCTCNet
def _raw_ctcnet_pack(model):
    C = model.layers[0][0]
    f0, f2 = model.feature_layer[0], model.feature_layer[2]
    w2 = f2.weight.detach().float()[:, :, 0]
    return {"f0w": f0.weight.detach().float()[:, 0, :].contiguous(), "f0b": f0.bias.detach().float().contiguous(),
            "f2w": FP._bf16(w2), "f2b": f2.bias.detach().float().contiguous(), "f2wt": FP._bf16(w2.t()),
            "stack": Stack([model.input_block] + list(model.convolutions),
                           [model.input_skip_bottleneck] + list(model.bottlenecks)),
            "head": FP.pack_head(model.output_block, C, bwd=True)}


class _RawCTCNetTrain(torch.autograd.Function):
    """RawCTCNet.forward (reference raw_ctcnet.py:117-153, positions=False) with its backward on the tensor-core
    kernels: featuriser Conv1d(1,F,fk,pad fk-1) + LeakyReLU (bandwidth kernel), 1x1 + LeakyReLU (dense), residual
    stack, head.  The output is fk-1 frames longer than the input (raw_ctcnet.py:58)."""

    @staticmethod
    def forward(ctx, model, pk, seq, *params):
        C, F, fk = model.layers[0][0], model.num_features, model.feature_kwidth
        B, _, T = seq.shape
        To = T + fk - 1
        f = torch.empty((B, To, F), dtype=torch.bfloat16, device=seq.device)
        _lib.call("wnb200_featurize_nlc", ops._dt(seq), B, T, F, fk, ops._p(seq), ops._p(pk["f0w"]), ops._p(pk["f0b"]),
                  _lib.ACT_BF16, ops._p(f), ops._stream())
        h0 = FP.dense(f, [0], pk["f2w"], pk["f2b"], F, leaky=1)
        skips = torch.empty((B, To, C), dtype=torch.float32, device=seq.device)
        saved, sk_act = stack_forward(h0, pk["stack"], skips)
        out, skips_act, h1 = head_forward(skips, pk["head"], seq.dtype, model.softmax, sk_act)
        del skips
        ctx.model, ctx.pk, ctx.params = model, pk, params
        ctx.keep = (seq, f, saved, skips_act, h1, out if model.softmax else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        with zero_pool(dout.device, _pool_floats(ctx.pk["stack"])):
            return _RawCTCNetTrain._backward(ctx, dout)

    @staticmethod
    def _backward(ctx, dout):
        model, pk, params = ctx.model, ctx.pk, ctx.params
        seq, f, saved, skips_act, h1, out = ctx.keep
        ctx.keep = None
        F, fk = model.num_features, model.feature_kwidth
        B, _, T = seq.shape
        dev = seq.device
        h0 = saved[0][0]
        dskips, ghead = head_backward(dout.contiguous(), out, model.softmax, pk["head"], skips_act, h1)
        dh0, gl, _ = stack_backward(pk["stack"], saved, dskips, True)
        dpre2 = leaky_bwd(dh0, h0)                                       # through the second LeakyReLU
        dw2 = wgrad_rows(dpre2, f, 0, F, F).unsqueeze(2)
        db2 = colsum(dpre2)
        df = FP.dense(dpre2, [0], pk["f2wt"], _zeros(F, dev), F)
        dw0, db0 = torch.zeros((F, fk), dtype=torch.float32, device=dev), _zeros(F, dev)
        need_dx = ctx.needs_input_grad[2]
        dseq = torch.zeros((B, T), dtype=torch.float32, device=dev) if need_dx else None
        _lib.call("wnb200_featurize_bwd_nlc", ops._dt(seq), B, T, F, fk, ops._p(df), ops._p(f), ops._p(seq),
                  ops._p(pk["f0w"]), ops._p(dw0), ops._p(db0), ops._p(dseq), ops._stream())
        grads = [dw0.unsqueeze(1), db0, dw2, db2] + [g for layer in gl for g in layer] + ghead
        return (None, None, dseq.view(B, 1, T).to(seq.dtype) if need_dx else None) + _cast(grads, params)


def raw_ctcnet_train_eligible(model, seq):
    C, F = model.layers[0][0], model.num_features
    return (FP.tc_dtype_ok(seq) and seq.is_cuda and seq.dim() == 3 and seq.shape[1] == 1 and F == C
            and model.out_dim == C and C in (128, 256) and not model.positions and FP._stack_ok(C, model.layers)
            and model.input_kernel_size <= 3 and model.feature_kwidth <= 4 and seq.shape[0] > 0 and seq.shape[2] > 0)


def raw_ctcnet_forward_train(model, seq):
    ops.check_device()
    pk = FP._cached(model, "raw_ctcnet_train", lambda: _raw_ctcnet_pack(model))
    f0, f2 = model.feature_layer[0], model.feature_layer[2]
    params = [f0.weight, f0.bias, f2.weight, f2.bias] + pk["stack"].params() + _head_params(model.output_block)
    return _RawCTCNetTrain.apply(model, pk, seq.contiguous(), *params)
