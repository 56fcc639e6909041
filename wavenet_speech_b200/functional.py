"""Autograd-aware functional layer over the C-ABI kernels (generic NCL path).

Each torch.autograd.Function below runs its forward AND backward in libwnb200 kernels; torch is
used for buffer allocation and for the tiny parameter re-layouts (weight slabs per tap).
Offsets follow the reference exactly: see `tap_offsets`.
"""
import torch

from . import ops
from .ops import EPI_GATE, EPI_LEAKY, EPI_MU, EPI_NONE, PRE_LNRELU, Term


# --------------------------------------------------------------------------- offsets
def autopad(k, d):
    """Same rule as the reference's autopad (modules/conv_ops.py:104-116): ceil((k-1)*d / 2)."""
    total = (k - 1) * d
    return (total + 1) // 2


def tap_offsets(k, d, causal):
    """Frame offset read by tap j.  Causal (conv_ops.py:28,44): j*d-(k-1)*d.  Non-causal
    (conv_ops.py:62,79): j*d-autopad(k,d)."""
    pad = (k - 1) * d if causal else autopad(k, d)
    return [j * d - pad for j in range(k)]


# --------------------------------------------------------------------------- weight re-layouts
# The kernels want one contiguous [rows, C] slab per tap.  Re-laying the weights out on every call was most of what the
# generic path launched in inference (round 1's smoke: 93 libwnb200 kernels among 242 ATen copies / fills / cats): the
# layouts are cached per tensor and rebuilt when its (data_ptr, _version) changes, like the tensor-core packs
# (`fastpath.invalidate_packs` also clears this cache, for writes through `.data` that bypass the version counter).
import weakref

_LAYOUTS = {}      # id(tensor) -> (weakref to it, {(kind, dtype): (version key, layout)}); entries die with the tensor
                   # (a WeakKeyDictionary would compare tensors with ==, which is element-wise)


def clear_layout_cache():
    _LAYOUTS.clear()


def _layout_slot(owner):
    k = id(owner)
    ent = _LAYOUTS.get(k)
    if ent is None or ent[0]() is not owner:
        def _gone(ref, k=k):
            cur = _LAYOUTS.get(k)
            if cur is not None and cur[0] is ref:
                _LAYOUTS.pop(k, None)
        ent = (weakref.ref(owner, _gone), {})
        _LAYOUTS[k] = ent
    return ent[1]


def _cached_layout(key_tensors, kind, dtype, build):
    """build() cached on the identity and version of every tensor in key_tensors (the first one owns the entry)."""
    owner = key_tensors[0]
    if torch.is_grad_enabled() and any(t.requires_grad for t in key_tensors if t is not None):
        return build()                     # training: the weights change every step, caching would only hold memory
    ver = tuple((None if t is None else (t.data_ptr(), t._version, t.dtype)) for t in key_tensors)
    try:
        slot = _layout_slot(owner)
    except TypeError:                      # not weak-referenceable: no caching
        return build()
    hit = slot.get((kind, dtype))
    if hit is None or hit[0] != ver:
        hit = (ver, build())
        slot[(kind, dtype)] = hit
    return hit[1]


def _slabs(w, dtype):
    """[M, C, k] (or [M, C]) -> [k, M, C] contiguous in the compute dtype."""
    def build():
        ww = w.unsqueeze(2) if w.dim() == 2 else w
        return ww.detach().to(dtype).permute(2, 0, 1).contiguous()
    return _cached_layout([w], "slabs", dtype, build)


def _slabs_t(w, dtype):
    """[M, C, k] -> [k, C, M] contiguous (transposed slabs for the data gradient)."""
    def build():
        ww = w.unsqueeze(2) if w.dim() == 2 else w
        return ww.detach().to(dtype).permute(2, 1, 0).contiguous()
    return _cached_layout([w], "slabs_t", dtype, build)


def _pack_gate(wt, ws, bt, bs, dtype):
    """Interleave tanh / sigmoid filters per 64 output channels: rows [128i, 128i+64) = tanh
    channels 64i.., rows [128i+64, 128i+128) = sigmoid channels 64i.. (see wnb200.h, EPI_GATE)."""
    return _cached_layout([wt, ws, bt, bs], "gate", dtype, lambda: _pack_gate_build(wt, ws, bt, bs, dtype))


def _pack_gate_build(wt, ws, bt, bs, dtype):
    M, C, k = wt.shape
    Mp = (M + 63) // 64 * 64
    nb = Mp // 64

    def padw(w):
        w = w.detach().to(dtype)
        if Mp != M:
            w = torch.cat([w, w.new_zeros(Mp - M, C, k)], 0)
        return w.view(nb, 64, C, k)

    def padb(b):
        b = b.detach().float() if b is not None else torch.zeros(M, device=wt.device)
        if Mp != M:
            b = torch.cat([b, b.new_zeros(Mp - M)], 0)
        return b.view(nb, 64)

    wg = torch.stack([padw(wt), padw(ws)], 1).reshape(2 * Mp, C, k).permute(2, 0, 1).contiguous()
    bg = torch.stack([padb(bt), padb(bs)], 1).reshape(2 * Mp).contiguous()
    return wg, bg


def _f32(b):
    if b is None:
        return None
    if b.dtype == torch.float32 and b.is_contiguous():
        return b.detach()                  # no copy: the kernels read the parameter in place
    return _cached_layout([b], "f32", torch.float32, lambda: b.detach().float().contiguous())


# --------------------------------------------------------------------------- generic conv
class _TapsConv(torch.autograd.Function):
    """out = epi(bias + sum_i sum_j w_i[:,:,j] @ pre_i(x_i[.., t + off_ij]))."""

    @staticmethod
    def forward(ctx, cfg, bias, *tensors):
        n = len(tensors) // 2
        xs = [ops.time_major(t) for t in tensors[:n]]
        ws = tensors[n:]
        dtype = xs[0].dtype
        M = ws[0].shape[0]
        T_out = cfg["T_out"] if cfg.get("T_out") is not None else xs[0].shape[2]
        terms = []
        for i in range(n):
            slabs = _slabs(ws[i], dtype)
            for j, off in enumerate(cfg["offsets"][i]):
                terms.append(Term(xs[i], slabs[j], off, cfg["pre_acts"][i]))
        out = ops.taps_fwd(terms, _f32(bias), M, T_out, cfg["epilogue"])
        ctx.cfg = cfg
        ctx.n = n
        ctx.has_bias = bias is not None
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.save_for_backward(out if cfg["epilogue"] == EPI_LEAKY else None, *xs, *ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg, n = ctx.cfg, ctx.n
        saved = ctx.saved_tensors
        out, xs, ws = saved[0], saved[1:1 + n], saved[1 + n:]
        if dout is None:
            return (None, None) + (None,) * (2 * n)
        dout = dout.contiguous()
        if cfg["epilogue"] == EPI_LEAKY:
            dout = ops.leaky_bwd(dout, out)
        dtype = dout.dtype
        M, T_out = dout.shape[1], dout.shape[2]
        gx, gw = [], []
        for i in range(n):
            x, w = xs[i], ws[i]
            offs = cfg["offsets"][i]
            if ctx.needs_input_grad[2 + i]:
                wt = _slabs_t(w, dtype)
                terms = [Term(dout, wt[j], -off) for j, off in enumerate(offs)]
                dx = ops.taps_fwd(terms, None, x.shape[1], x.shape[2])
                if cfg["pre_acts"][i]:
                    dx = ops.leaky_bwd(dx, x)
                gx.append(dx)
            else:
                gx.append(None)
            if ctx.needs_input_grad[2 + n + i]:
                k = len(offs)
                dw = torch.zeros((k, M, x.shape[1]), dtype=torch.float32, device=dout.device)
                for j, off in enumerate(offs):
                    ops.taps_wgrad(x, off, cfg["pre_acts"][i], dout, dw[j])
                dw = dw.permute(1, 2, 0)
                if w.dim() == 2:
                    dw = dw[:, :, 0]
                gw.append(dw.to(w.dtype))
            else:
                gw.append(None)
        gb = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            gb = ops.channel_reduce(dout).to(ctx.bias_dtype)
        return (None, gb) + tuple(gx) + tuple(gw)


def multi_conv(terms, bias, epilogue=EPI_NONE, T_out=None):
    """terms: list of (x, weight[M,C,k] or [M,C], offsets(list of k ints), pre_act)."""
    cfg = {"offsets": [list(t[2]) for t in terms], "pre_acts": [int(t[3]) for t in terms],
           "epilogue": epilogue, "T_out": T_out}
    xs = [t[0] for t in terms]
    ws = [t[1] for t in terms]
    return _TapsConv.apply(cfg, bias, *xs, *ws)


def conv_taps(x, weight, bias, offsets, pre_act=0, epilogue=EPI_NONE, T_out=None):
    return multi_conv([(x, weight, offsets, pre_act)], bias, epilogue, T_out)


# --------------------------------------------------------------------------- residual block
class _ResBlock(torch.autograd.Function):
    """ResidualBlock.forward (reference modules/block.py:54-82) in three launches:
    gate = tanh(conv_t x) * sigmoid(conv_s x);  res = Wres gate + Wproj x + b;  skip = Wskip gate + b."""

    @staticmethod
    def forward(ctx, offsets, x, wt, bt, ws, bs, wres, bres, wskip, bskip, wproj, bproj):
        x = ops.time_major(x)
        dtype = x.dtype
        M, C, k = wt.shape
        T = x.shape[2]
        need_bwd = any(ctx.needs_input_grad)
        if k > ops.MAX_SRC:
            # wide kernels (the reference trains with widths up to 32: pretrain_tnt.py:98,120): the two pre-activations
            # are accumulated over chunks of MAX_SRC taps (chained launches), then gated by the stand-alone gate kernel
            wts, wss = _slabs(wt, dtype), _slabs(ws, dtype)
            a = ops.taps_fwd([Term(x, wts[j], offsets[j]) for j in range(k)], _f32(bt), M, T)
            b = ops.taps_fwd([Term(x, wss[j], offsets[j]) for j in range(k)], _f32(bs), M, T)
            if need_bwd:
                act, th, sg = ops.gate_fwd(a, b, want_parts=True)
            else:
                act, th, sg = ops.gate_fwd(a, b), None, None
            del a, b
        else:
            wg, bg = _pack_gate(wt, ws, bt, bs, dtype)
            gterms = [Term(x, wg[j], offsets[j]) for j in range(k)]
            if need_bwd:
                act, th, sg = ops.taps_fwd(gterms, bg, M, T, EPI_GATE, want_gate_parts=True)
            else:
                act, th, sg = ops.taps_fwd(gterms, bg, M, T, EPI_GATE), None, None
        wres2, wproj2, wskip2 = _slabs(wres, dtype)[0], _slabs(wproj, dtype)[0], _slabs(wskip, dtype)[0]
        b_res = _cached_layout([bres, bproj], "bias_sum", torch.float32,
                               lambda: (bres.detach().float() + bproj.detach().float()).contiguous())
        res = ops.taps_fwd([Term(act, wres2), Term(x, wproj2)], b_res, M, T)
        skip = ops.taps_fwd([Term(act, wskip2)], _f32(bskip), M, T)
        ctx.offsets = list(offsets)
        ctx.save_for_backward(x, act, th, sg, wt, ws, wres, wskip, wproj)
        ctx.set_materialize_grads(False)
        ctx.bias_dtypes = (bt.dtype, bs.dtype, bres.dtype, bskip.dtype, bproj.dtype)
        return res, skip

    @staticmethod
    def backward(ctx, dres, dskip):
        x, act, th, sg, wt, ws, wres, wskip, wproj = ctx.saved_tensors
        offs = ctx.offsets
        nothing = (None,) * 12
        if dres is None and dskip is None:
            return nothing
        dtype = x.dtype
        M, C, k = wt.shape
        B, _, T = x.shape
        dev = x.device
        if dres is not None:
            dres = dres.contiguous()
        if dskip is not None:
            dskip = dskip.contiguous()
        # d(gate)
        terms = []
        if dres is not None:
            terms.append(Term(dres, _slabs_t(wres, dtype)[0]))
        if dskip is not None:
            terms.append(Term(dskip, _slabs_t(wskip, dtype)[0]))
        dact = ops.taps_fwd(terms, None, M, T)
        dab = ops.gate_bwd(dact, th, sg)                        # [B, 2M, T] = (d tanh-pre ; d sigmoid-pre)
        # d(x)
        wab_t = torch.cat([_slabs_t(wt, dtype), _slabs_t(ws, dtype)], 2)   # [k, C, 2M]
        terms = [Term(dab, wab_t[j].contiguous(), -offs[j]) for j in range(k)]
        if dres is not None:
            terms.append(Term(dres, _slabs_t(wproj, dtype)[0]))
        dx = ops.taps_fwd(terms, None, C, T) if ctx.needs_input_grad[1] else None
        # parameter gradients (fp32 accumulation)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        dwab = z(k, 2 * M, C)
        for j in range(k):
            ops.taps_wgrad(x, offs[j], 0, dab, dwab[j])
        dwt = dwab[:, :M].permute(1, 2, 0).to(wt.dtype)
        dws = dwab[:, M:].permute(1, 2, 0).to(ws.dtype)
        dbab = ops.channel_reduce(dab)
        bd = ctx.bias_dtypes
        dbt, dbs = dbab[:M].to(bd[0]), dbab[M:].to(bd[1])
        dwres = dwproj = dbres = dbproj = dwskip = dbskip = None
        if dres is not None:
            dwres = ops.taps_wgrad(act, 0, 0, dres, z(M, M)).view(wres.shape).to(wres.dtype)
            dwproj = ops.taps_wgrad(x, 0, 0, dres, z(M, C)).view(wproj.shape).to(wproj.dtype)
            db = ops.channel_reduce(dres)
            dbres, dbproj = db.to(bd[2]), db.to(bd[4])
        if dskip is not None:
            dwskip = ops.taps_wgrad(act, 0, 0, dskip, z(M, M)).view(wskip.shape).to(wskip.dtype)
            dbskip = ops.channel_reduce(dskip).to(bd[3])
        return (None, dx, dwt, dbt, dws, dbs, dwres, dbres, dwskip, dbskip, dwproj, dbproj)


def residual_block(x, p, offsets):
    """p: object with conv_tanh / conv_sigmoid / conv1x1_residual / conv1x1_skip / residual_proj."""
    return _ResBlock.apply(list(offsets), x,
                           p.conv_tanh.conv1d.weight, p.conv_tanh.conv1d.bias,
                           p.conv_sigmoid.conv1d.weight, p.conv_sigmoid.conv1d.bias,
                           p.conv1x1_residual.weight, p.conv1x1_residual.bias,
                           p.conv1x1_skip.weight, p.conv1x1_skip.bias,
                           p.residual_proj.weight, p.residual_proj.bias)


class _SkipAccum(torch.autograd.Function):
    """skips += Wbn skip + bbn, in place (reference wavenet.py:100 allocates a new tensor per layer)."""

    @staticmethod
    def forward(ctx, skips, skip, w, b):
        skip = ops.time_major(skip)
        dtype = skip.dtype
        M = w.shape[0]
        ops.taps_fwd([Term(skip, _slabs(w, dtype)[0])], _f32(b), M, skip.shape[2], EPI_NONE, out=skips,
                     accumulate=True)
        ctx.mark_dirty(skips)
        ctx.save_for_backward(skip, w)
        ctx.bias_dtype = b.dtype
        return skips

    @staticmethod
    def backward(ctx, dout):
        skip, w = ctx.saved_tensors
        dout = dout.contiguous()
        dtype = dout.dtype
        M, C = w.shape[0], w.shape[1]
        dskip = ops.taps_fwd([Term(dout, _slabs_t(w, dtype)[0])], None, C, skip.shape[2])
        dw = ops.taps_wgrad(skip, 0, 0, dout, torch.zeros((M, C), dtype=torch.float32, device=dout.device))
        db = ops.channel_reduce(dout).to(ctx.bias_dtype)
        return dout, dskip, dw.view(w.shape).to(w.dtype), db


def skip_accumulate(skips, skip, w, b):
    """First call (skips is None) starts the running sum; later calls accumulate in place."""
    if skips is None:
        return conv_taps(skip, w, b, [0])
    return _SkipAccum.apply(skips, skip, w, b)


def output_stack(x, w1, b1, w3, b3):
    """LeakyReLU -> 1x1 -> LeakyReLU -> 1x1 (reference wavenet.py:67-71): two launches, the
    activations ride on the load of the first and the epilogue of the first."""
    h = conv_taps(x, w1, b1, [0], pre_act=1, epilogue=EPI_LEAKY)
    return conv_taps(h, w3, b3, [0])


# --------------------------------------------------------------------------- softmax / pool / LN / loss
class _ChannelSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, log_mode):
        y = ops.softmax_fwd(x, log_mode)
        ctx.log_mode = log_mode
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.softmax_bwd(y, dy, ctx.log_mode), None


def channel_softmax(x, log=False):
    return _ChannelSoftmax.apply(x, bool(log))


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pool):
        ctx.T, ctx.pool = x.shape[2], pool
        return ops.avgpool_fwd(x, pool)

    @staticmethod
    def backward(ctx, dy):
        return ops.avgpool_bwd(dy, ctx.T, ctx.pool), None


def avg_pool(x, pool):
    return _AvgPool.apply(x, int(pool))


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        g = gamma.detach().float().reshape(-1).contiguous()
        b = beta.detach().float().reshape(-1).contiguous()
        y, stats = ops.layernorm_fwd(x, g, b, eps)
        ctx.eps = eps
        ctx.save_for_backward(x, g, stats, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, stats, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.layernorm_bwd(x, g, stats, ctx.eps, dy)
        dgamma, dbeta = ops.layernorm_bwd_params(x, stats, dy)
        return dx, dgamma.view(gamma.shape).to(gamma.dtype), dbeta.view(gamma.shape).to(gamma.dtype), None


def layer_norm(x, gamma, beta, eps=1e-6):
    return _LayerNorm.apply(x, gamma, beta, float(eps))


class _XentSum(torch.autograd.Function):
    """sum_{b,t} -log softmax(logits[b,:,t])[target[b,t]] in one pass (legacy_code/train.py:36-39
    loops over t in Python and launches T-1 CrossEntropyLoss ops)."""

    @staticmethod
    def forward(ctx, logits, target):
        logits = logits.contiguous()
        loss_bt, lse = ops.xent_fwd(logits, target)
        ctx.save_for_backward(logits, target, lse)
        return ops.sum_f32(loss_bt)

    @staticmethod
    def backward(ctx, g):
        logits, target, lse = ctx.saved_tensors
        gs = g.detach().float().reshape(1).contiguous()
        return ops.xent_bwd(logits, target, lse, gs), None


def cross_entropy_sum(logits, target):
    return _XentSum.apply(logits, target)


class _Gate(torch.autograd.Function):
    """Stand-alone tanh(a) * sigmoid(b) (reference block.py:184-185)."""

    @staticmethod
    def forward(ctx, a, b):
        out, th, sg = ops.gate_fwd(a, b, want_parts=True)
        ctx.save_for_backward(th, sg)
        return out

    @staticmethod
    def backward(ctx, g):
        th, sg = ctx.saved_tensors
        shape = g.shape
        g3 = g.contiguous().view(1, -1, 1) if g.dim() != 3 else g.contiguous()
        dab = ops.gate_bwd(g3, th.view_as(g3), sg.view_as(g3))
        c = g3.shape[1]
        return dab[:, :c].reshape(shape), dab[:, c:].reshape(shape)


def gated_activation(a, b):
    return _Gate.apply(a, b)


class _MUGate(torch.autograd.Function):
    """g1 * tanh(g2 * h + g3 * u) from the four stacked pre-activations (reference block.py:213-220)."""

    @staticmethod
    def forward(ctx, pre, h):
        from . import _lib
        pre, h = pre.contiguous(), h.contiguous()
        B, C, T = h.shape
        out = torch.empty_like(h)
        _lib.call("wnb200_mu_gate_fwd", ops._dt(h), B, C, T, ops._p(pre), ops._p(h), ops._p(out), ops._stream())
        ctx.save_for_backward(pre, h)
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import _lib
        pre, h = ctx.saved_tensors
        B, C, T = h.shape
        dpre, dh = torch.empty_like(pre), torch.empty_like(h)
        _lib.call("wnb200_mu_gate_bwd", ops._dt(h), B, C, T, ops._p(pre), ops._p(h), ops._p(dout.contiguous()),
                  ops._p(dpre), ops._p(dh), ops._stream())
        return dpre, dh


def multiplicative_unit(h, convs, offsets):
    """MultiplicativeUnit.forward: ONE tap-sum launch for the four causal convolutions (their filters stacked on
    the output-channel axis) + one gate launch.  convs = (gate1, gate2, gate3, update) nn.Conv1d holders."""
    if len(offsets) <= ops.MAX_SRC and not needs_grad(h, *[q for c in convs for q in (c.weight, c.bias)]):
        return multiplicative_unit_fused(h, convs, offsets)       # inference: the gate is the contraction's epilogue
    w = torch.cat([c.weight for c in convs], 0)
    b = torch.cat([c.bias for c in convs], 0)
    pre = conv_taps(h, w, b, offsets)
    return _MUGate.apply(pre, h)


# --------------------------------------------------------------------------- ByteNet blocks (block.py:86-173)
def _flat32(p):
    return _cached_layout([p], "flat32", torch.float32, lambda: p.detach().float().reshape(-1).contiguous())


class _FusedConv(torch.autograd.Function):
    """out = bias + sum_j W[:, :, j] @ pre(x)[.., t + off_j]  (+ residual), with pre = ReLU(LayerNorm(.)) applied AS THE
    OPERAND IS LOADED when gamma / beta are given: the normalised tensor of the reference's
    [LayerNorm, ReLU, conv] triples (block.py:103-105, 108-110, 150-160) is never stored.  Launches: statistics +
    contraction; backward: transposed contraction, LayerNorm+ReLU backward (dx, dgamma, dbeta in one pass), weight
    gradients with the same on-the-fly operand."""

    @staticmethod
    def forward(ctx, cfg, x, gamma, beta, weight, bias, residual):
        x = x.contiguous()
        dtype = x.dtype
        M = weight.shape[0]
        T = x.shape[2]
        offs = cfg["offsets"]
        ln = None
        if gamma is not None:
            ln = (ops.ln_stats(x, cfg["eps"]), _flat32(gamma), _flat32(beta))
        slabs = _slabs(weight, dtype)
        pre = PRE_LNRELU if ln is not None else 0
        terms = [Term(x, slabs[j], off, pre, ln) for j, off in enumerate(offs)]
        res = None if residual is None else residual.contiguous()
        out = ops.taps_fwd(terms, _f32(bias), M, T, residual=res)
        ctx.cfg = cfg
        ctx.bias_dtype = None if bias is None else bias.dtype
        ctx.save_for_backward(x, weight, gamma, beta, None if ln is None else ln[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg = ctx.cfg
        x, weight, gamma, beta, stats = ctx.saved_tensors
        offs = cfg["offsets"]
        dout = dout.contiguous()
        dtype = dout.dtype
        M, T = dout.shape[1], dout.shape[2]
        C = x.shape[1]
        ln = None
        if gamma is not None:
            ln = (stats, _flat32(gamma), _flat32(beta))
        need = ctx.needs_input_grad          # (cfg, x, gamma, beta, weight, bias, residual)
        dx = dgamma = dbeta = dw = db = None
        if need[1] or need[2] or need[3]:
            wt = _slabs_t(weight, dtype)
            dh = ops.taps_fwd([Term(dout, wt[j], -off) for j, off in enumerate(offs)], None, C, T)
            if ln is None:
                dx = dh
            else:
                dx, dg, dbt = ops.ln_relu_bwd(x, ln[0], ln[1], ln[2], cfg["eps"], dh, want_dx=need[1])
                dgamma = dg.view(gamma.shape).to(gamma.dtype)
                dbeta = dbt.view(gamma.shape).to(gamma.dtype)
        if need[4]:
            k = len(offs)
            dwk = torch.zeros((k, M, C), dtype=torch.float32, device=dout.device)
            for j, off in enumerate(offs):
                ops.taps_wgrad(x, off, PRE_LNRELU if ln is not None else 0, dout, dwk[j], ln=ln)
            dw = dwk.permute(1, 2, 0)
            if weight.dim() == 2:
                dw = dw[:, :, 0]
            dw = dw.to(weight.dtype)
        if ctx.bias_dtype is not None and need[5]:
            db = ops.channel_reduce(dout).to(ctx.bias_dtype)
        return None, dx, dgamma, dbeta, dw, db, (dout if need[6] else None)


def fused_conv(x, weight, bias, offsets, ln=None, residual=None):
    """conv over `offsets` of ReLU(LayerNorm(x)) (ln = the LayerNorm module) or of x, plus an optional residual."""
    gamma = beta = None
    cfg = {"offsets": [int(o) for o in offsets], "eps": 0.0}
    if ln is not None:
        if ln.dim != 1:
            raise NotImplementedError("LayerNorm kernel normalises dim=1 of a (B, C, T) tensor")
        gamma, beta = ln.gamma, ln.beta
        cfg["eps"] = float(ln.eps)
    return _FusedConv.apply(cfg, x, gamma, beta, weight, bias, residual)


class _LNReLU(torch.autograd.Function):
    """ReLU(LayerNorm(x)) stored (one statistics + one apply launch); the training form of the MU block keeps it for the
    MultiplicativeUnit's backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        g, b = _flat32(gamma), _flat32(beta)
        stats = ops.ln_stats(x, eps)
        ctx.eps = eps
        ctx.save_for_backward(x, stats, gamma, beta)
        return ops.ln_relu_fwd(x, stats, g, b)

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma, beta = ctx.saved_tensors
        dx, dg, db = ops.ln_relu_bwd(x, stats, _flat32(gamma), _flat32(beta), ctx.eps, dy)
        return dx, dg.view(gamma.shape).to(gamma.dtype), db.view(gamma.shape).to(gamma.dtype), None


def ln_relu(x, ln):
    if ln.dim != 1:
        raise NotImplementedError("LayerNorm kernel normalises dim=1 of a (B, C, T) tensor")
    return _LNReLU.apply(x, ln.gamma, ln.beta, float(ln.eps))


def _pack_mu(convs, dtype):
    """The four convolutions of a MultiplicativeUnit in EPI_MU's row order (wnb200.h): per 32 channels, rows [0, 64) =
    channels 0..15 x (gate1, gate2, gate3, update), rows [64, 128) = channels 16..31 likewise.  -> ([k, rows, C], [rows])"""
    ws = [c.weight for c in convs]
    bs = [c.bias for c in convs]

    def build():
        M, C, k = ws[0].shape
        nt = (M + 31) // 32
        r = torch.arange(nt * 128, device=ws[0].device)
        tile, rr = r // 128, r % 128
        ch = tile * 32 + torch.where(rr < 64, rr // 4, 16 + (rr - 64) // 4)
        unit = rr % 4
        ok = ch < M
        src = torch.where(ok, unit * M + ch, torch.full_like(ch, 4 * M))       # row 4M = the zero row
        w = torch.cat([w_.detach().to(dtype) for w_ in ws] + [ws[0].new_zeros((1, C, k), dtype=dtype)], 0)
        b = torch.cat([b_.detach().float() for b_ in bs] + [bs[0].new_zeros(1, dtype=torch.float32)], 0)
        return w[src].permute(2, 0, 1).contiguous(), b[src].contiguous()
    return _cached_layout(ws + bs, "mu", dtype, build)


def multiplicative_unit_fused(h, convs, offsets, ln=None, residual=None):
    """MultiplicativeUnit.forward in ONE launch (no autograd): the four convolutions' contraction with the gate
    g1 * tanh(g2 * h + g3 * tanh(u)) as its epilogue.  ln = a LayerNorm module: h is ReLU(LayerNorm(h)) formed on the
    fly, for the operand AND for the gate's h."""
    h = h.contiguous()
    wk, bk = _pack_mu(convs, h.dtype)
    lnp = None
    if ln is not None:
        lnp = (ops.ln_stats(h, float(ln.eps)), _flat32(ln.gamma), _flat32(ln.beta))
    terms = [Term(h, wk[j], off, PRE_LNRELU if lnp is not None else 0, lnp) for j, off in enumerate(offsets)]
    return ops.taps_fwd(terms, bk, h.shape[1], h.shape[2], EPI_MU, mu_h=h, mu_h_ln=lnp is not None, residual=residual)


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _Positions(torch.autograd.Function):
    """out + hardtanh(w * t + b) (reference raw_ctcnet.py:131-135), with the gradients of x AND of the position layer's
    two parameters (the reference trains RawCTCNet(positions=True): pretrain_tnt.py:121-124, tests/kmer_stay_prediction.py:52)."""

    @staticmethod
    def forward(ctx, x, w, b, t0):
        out = x.contiguous().clone()
        wf = w.detach().float().reshape(-1).contiguous()
        bf = b.detach().float().reshape(-1).contiguous()
        ops.positions_add_(out, wf, bf, t0)
        ctx.save_for_backward(wf, bf)
        ctx.t0, ctx.wshape, ctx.wdt, ctx.bdt = t0, w.shape, w.dtype, b.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        wf, bf = ctx.saved_tensors
        dw = db = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            gc = g.contiguous()
            B, F, T = gc.shape
            dw = torch.zeros(F, dtype=torch.float32, device=g.device)
            db = torch.zeros(F, dtype=torch.float32, device=g.device)
            _lib.call("wnb200_positions_bwd", ops._dt(gc), B, F, T, int(ctx.t0), ops._p(wf), ops._p(bf), ops._p(gc),
                      ops._p(dw), ops._p(db), ops._stream())
            dw, db = dw.view(ctx.wshape).to(ctx.wdt), db.to(ctx.bdt)
        return g, dw, db, None


def positions_mix(x, w, b, t0=0):
    return _Positions.apply(x, w, b, int(t0))


# --------------------------------------------------------------------------- CTC
class _CTCSum(torch.autograd.Function):
    """sum_b -log p(labels_b | act_b) with the blank at class 0 and pre-softmax activations -- the
    warpctc_pytorch.CTCLoss convention the reference trains with (legacy_code/train.py:42-46).  `layout`:
    "tbc" = (T, B, C) as the reference passes it, "bct" = the classifier's (B, C, T) output, read in place."""

    @staticmethod
    def forward(ctx, act, labels, label_lengths, act_lengths, layout):
        from . import _lib
        ops._need_cuda(act)
        ops.check_device()
        if layout == "tbc":
            T, B, L = act.shape
            st, sb, sc = act.stride()
        else:
            B, L, T = act.shape
            sb, sc, st = act.stride()
        dev = act.device
        ll = torch.as_tensor(label_lengths, dtype=torch.int64).cpu()
        assert ll.numel() == B, "ctc: one label length per read"
        max_len = int(ll.max()) if B > 0 else 0
        offs = torch.zeros(B + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(ll, 0)
        labels = labels.to(device=dev, dtype=torch.int32).contiguous()
        assert labels.numel() == int(offs[-1]), "ctc: labels must be the concatenation of the reads' labels"
        offs = offs.to(dev, non_blocking=True)
        al = None
        if act_lengths is not None:
            al = torch.as_tensor(act_lengths, dtype=torch.int32).to(dev, non_blocking=True).contiguous()
        nbytes = _lib.load().wnb200_ctc_workspace_bytes(B, L, T, max_len)
        ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.call("wnb200_ctc_fwd", ops._dt(act), B, L, T, max_len, ops._p(act), sb, sc, st, ops._p(labels),
                  ops._p(offs), ops._p(al), ops._p(ws), ops._p(nll), ops._stream())
        ctx.dims = (B, L, T, max_len, sb, sc, st)
        ctx.save_for_backward(act, labels, offs, al, ws, nll)
        return ops.sum_f32(nll)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        act, labels, offs, al, ws, nll = ctx.saved_tensors
        B, L, T, max_len, sb, sc, st = ctx.dims
        grad = torch.empty_strided(act.shape, act.stride(), dtype=act.dtype, device=act.device)
        gs = g.detach().float().reshape(1).contiguous()
        _lib.call("wnb200_ctc_bwd", ops._dt(act), B, L, T, max_len, ops._p(labels), ops._p(offs), ops._p(al),
                  ops._p(ws), ops._p(nll), ops._p(gs), ops._p(grad), sb, sc, st, ops._stream())
        return grad, None, None, None, None


def ctc_loss_sum(act, labels, label_lengths, act_lengths=None, layout="bct"):
    """CTC negative log-likelihood summed over the batch (blank = 0, softmax applied inside)."""
    assert layout in ("bct", "tbc")
    return _CTCSum.apply(act, labels, label_lengths, act_lengths, layout)


class CTCLoss(torch.nn.Module):
    """Call-compatible stand-in for warpctc_pytorch.CTCLoss as the reference uses it
    (legacy_code/train.py:116,46: `ctc_loss_fn(probs, labels, prob_lengths, lengths)` with probs (T, B, C)
    pre-softmax, labels a flat int tensor with 0 = blank, result summed over the batch, shape (1,))."""

    def forward(self, acts, labels, act_lens, label_lens):
        return ctc_loss_sum(acts, labels, label_lens, act_lens, layout="tbc").reshape(1)
