"""Tensor-level wrappers over the C-ABI (no autograd here; see functional.py).

Every function takes CUDA tensors, allocates outputs with torch (device memory plumbing only) and
enqueues the kernels on torch's current stream.  A CPU tensor raises: there is no CPU path.
"""
import ctypes

import torch

from . import _lib

EPI_NONE, EPI_LEAKY, EPI_GATE, EPI_MU = 0, 1, 2, 3
PRE_NONE, PRE_LEAKY, PRE_LNRELU = 0, 1, 2
MAX_SRC = 4
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("wavenet_speech_b200 supports float32 and bfloat16 tensors, got %s" % t.dtype)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("wavenet_speech_b200 runs on CUDA (sm_100a) only; got a %s tensor. "
                               "There is no CPU fallback." % t.device)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def device_guard(fn):
    """Run a module's forward with the INPUT tensor's device current: streams, the device check and every launch then
    address that device even if the caller's current device is another one (one process driving several GPUs)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, x, *a, **k):
        if torch.is_tensor(x) and x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):
                return fn(self, x, *a, **k)
        return fn(self, x, *a, **k)
    return wrapper


_checked = {}


def check_device():
    dev = torch.cuda.current_device()
    if dev not in _checked:
        _lib.call("wnb200_check_device")
        _checked[dev] = True


def time_major(x):
    """Make the time stride 1 (the only layout requirement of the NCL kernels)."""
    return x if x.stride(2) == 1 and x.dim() == 3 else x.contiguous()


class Term(object):
    """One (x, weight-slab, offset) term.  x: [B, C, T_src] with stride(2)==1; w: [rows, C] contiguous.
    pre_act == PRE_LNRELU: `ln` = (stats [B, T_src, 2], gamma [C], beta [C]) fp32, the source is read as
    ReLU(LayerNorm(x)) without that tensor being stored."""
    __slots__ = ("x", "w", "t_off", "pre_act", "ln")

    def __init__(self, x, w, t_off=0, pre_act=0, ln=None):
        self.x, self.w, self.t_off, self.pre_act, self.ln = x, w, int(t_off), int(pre_act), ln
        assert (self.pre_act == PRE_LNRELU) == (ln is not None)


def _fill_ln(dst, term):
    stats, gamma, beta = term.ln
    assert stats.dtype == gamma.dtype == beta.dtype == torch.float32
    assert stats.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous()
    assert stats.shape == (term.x.shape[0], term.x.shape[2], 2) and gamma.numel() == term.x.shape[1]
    dst.stats, dst.gamma, dst.beta = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()


def _fill(src, term):
    x = term.x
    src.x = x.data_ptr()
    src.w = 0 if term.w is None else term.w.data_ptr()
    src.batch_stride = x.stride(0)
    src.chan_stride = x.stride(1)
    src.C = x.shape[1]
    src.T_src = x.shape[2]
    src.t_off = term.t_off
    src.pre_act = term.pre_act


def taps_fwd(terms, bias, M, T_out, epilogue=EPI_NONE, out=None, accumulate=False, want_gate_parts=False,
             residual=None, mu_h=None, mu_h_ln=False):
    """out[b,m,t] (+)= epi(bias[m] + sum_terms w @ pre(x[.., t+off])) [+ residual].  Returns out (and th, sg).
    epilogue == EPI_MU: `mu_h` is the h of the MultiplicativeUnit's gate (the weights' rows in the MU packing)."""
    x0 = terms[0].x
    _need_cuda(x0)
    check_device()
    B = x0.shape[0]
    dt = _dt(x0)
    for tm in terms:
        assert tm.x.dtype == x0.dtype and tm.w.dtype == x0.dtype, "mixed dtypes in taps_fwd"
        assert tm.x.stride(2) == 1 and tm.w.is_contiguous()
        assert tm.w.shape[1] == tm.x.shape[1]
    if out is None:
        out = torch.empty((B, M, T_out), dtype=x0.dtype, device=x0.device)
        assert not accumulate
    th = sg = None
    if want_gate_parts:
        th = torch.empty_like(out)
        sg = torch.empty_like(out)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    # more than MAX_SRC terms: chain launches through the accumulate path (epilogue applied last)
    chunks = [terms[i:i + MAX_SRC] for i in range(0, len(terms), MAX_SRC)]
    late_leaky = len(chunks) > 1 and epilogue == EPI_LEAKY      # wide kernels (RawCTCNet featuriser, fk up to 64)
    if len(chunks) > 1:
        assert epilogue in (EPI_NONE, EPI_LEAKY) and not (late_leaky and accumulate), \
            "more than %d terms: the gate epilogue is applied by the caller (functional._ResBlock)" % MAX_SRC
    extended = residual is not None or epilogue == EPI_MU or any(tm.ln is not None for tm in terms)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == out.dtype and residual.is_contiguous()
        assert epilogue != EPI_GATE and not late_leaky
    if epilogue == EPI_MU:
        assert len(chunks) == 1 and mu_h is not None and mu_h.dtype == x0.dtype
        assert mu_h is terms[0].x if mu_h_ln else (mu_h.shape == out.shape and mu_h.is_contiguous())
    for ci, chunk in enumerate(chunks):
        arr = (_lib.Src * len(chunk))()
        for i, tm in enumerate(chunk):
            _fill(arr[i], tm)
        if not extended:
            _lib.call("wnb200_taps_fwd", dt, B, T_out, M, len(chunk), arr, _p(bias if ci == 0 else None),
                      EPI_NONE if late_leaky else epilogue, 1 if (accumulate or ci > 0) else 0, _p(out), _p(th), _p(sg),
                      _stream())
            continue
        lns = (_lib.Ln * len(chunk))()
        for i, tm in enumerate(chunk):
            if tm.ln is not None:
                _fill_ln(lns[i], tm)
        a = _lib.Taps()
        a.dtype, a.B, a.T_out, a.M, a.nsrc = dt, B, T_out, M, len(chunk)
        a.epilogue = EPI_NONE if late_leaky else epilogue
        a.accumulate = 1 if (accumulate or ci > 0) else 0
        a.srcs, a.ln = arr, lns
        a.bias = 0 if (bias is None or ci > 0) else bias.data_ptr()
        a.out = out.data_ptr()
        a.th = 0 if th is None else th.data_ptr()
        a.sg = 0 if sg is None else sg.data_ptr()
        a.residual = residual.data_ptr() if (residual is not None and ci == len(chunks) - 1) else 0
        a.mu_h = 0 if mu_h is None else mu_h.data_ptr()
        a.mu_h_ln = 1 if mu_h_ln else 0
        _lib.call("wnb200_taps_fwd_ex", ctypes.byref(a), _stream())
    if late_leaky:       # x * (x > 0 ? 1 : 0.01), in place, with the LeakyReLU-backward kernel applied to (x, x)
        _lib.call("wnb200_leaky_bwd", dt, out.numel(), _p(out), _p(out), _p(out), _stream())
    if want_gate_parts:
        return out, th, sg
    return out


def taps_wgrad(x, t_off, pre_act, dout, dw, ln=None):
    """dw[m,c] += sum_{b,t} dout[b,m,t] * pre(x[b,c,t+t_off]); dw fp32 [M, C]."""
    _need_cuda(x, dout, dw)
    assert dout.is_contiguous() and dw.dtype == torch.float32 and dw.is_contiguous()
    assert x.stride(2) == 1 and x.dtype == dout.dtype
    B, M, T_out = dout.shape
    src = _lib.Src()
    tm = Term(x, None, t_off, pre_act, ln)
    _fill(src, tm)
    if ln is None:
        _lib.call("wnb200_taps_wgrad", _dt(x), B, T_out, M, ctypes.byref(src), _p(dout), _p(dw), _stream())
    else:
        lnc = _lib.Ln()
        _fill_ln(lnc, tm)
        _lib.call("wnb200_taps_wgrad_ex", _dt(x), B, T_out, M, ctypes.byref(src), ctypes.byref(lnc), _p(dout), _p(dw),
                  _stream())
    return dw


def ln_stats(x, eps):
    """(mean over channels, 1 / (unbiased std + eps)) per frame of a contiguous (B, C, T) tensor -> fp32 (B, T, 2)."""
    _need_cuda(x)
    check_device()
    assert x.is_contiguous() and x.dim() == 3
    B, C, T = x.shape
    stats = torch.empty((B, T, 2), dtype=torch.float32, device=x.device)
    _lib.call("wnb200_ln_stats", _dt(x), B, C, T, _p(x), float(eps), _p(stats), _stream())
    return stats


def ln_relu_fwd(x, stats, gamma, beta):
    B, C, T = x.shape
    y = torch.empty_like(x)
    _lib.call("wnb200_ln_relu_fwd", _dt(x), B, C, T, _p(x), _p(stats), _p(gamma), _p(beta), _p(y), _stream())
    return y


def ln_relu_bwd(x, stats, gamma, beta, eps, dy, want_dx=True):
    """Backward of ReLU(LayerNorm(x)) -> (dx or None, dgamma fp32 [C], dbeta fp32 [C])."""
    B, C, T = x.shape
    dy = dy.contiguous()
    assert dy.dtype == x.dtype and dy.shape == x.shape
    dx = torch.empty_like(x) if want_dx else None
    dgamma = torch.zeros(C, dtype=torch.float32, device=x.device)
    dbeta = torch.zeros(C, dtype=torch.float32, device=x.device)
    _lib.call("wnb200_ln_relu_bwd", _dt(x), B, C, T, _p(x), _p(stats), _p(gamma), _p(beta), float(eps), _p(dy), _p(dx),
              _p(dgamma), _p(dbeta), _stream())
    return dx, dgamma, dbeta


def linear_frame(frame, weight, bias, dilation):
    """LinearConv1d.linear: frame (N, Cin, rf), weight (Cout, Cin, k) -> (N, Cout)."""
    _need_cuda(frame, weight)
    check_device()
    frame = frame.contiguous()
    N, Cin, rf = frame.shape
    Cout, _, k = weight.shape
    assert weight.is_contiguous() and weight.dtype == frame.dtype and rf == k + (dilation - 1) * (k - 1)
    y = torch.empty((N, Cout), dtype=frame.dtype, device=frame.device)
    _lib.call("wnb200_linear_frame", _dt(frame), N, Cin, Cout, k, int(dilation), _p(weight), _p(bias), _p(frame), _p(y),
              _stream())
    return y


def linear_step(x, weight, bias, dilation, step, hist, out=None):
    """One frame of the incremental evaluation: x (N, Cin) is frame number `step`; hist (rf, N, Cin) is the ring the
    call reads the earlier taps from and files x into.  -> (N, Cout)."""
    _need_cuda(x, weight)
    check_device()
    N, Cin = x.shape
    Cout, _, k = weight.shape
    assert x.is_contiguous() and weight.is_contiguous() and weight.dtype == x.dtype
    if k > 1:
        assert hist.shape == ((k - 1) * dilation + 1, N, Cin) and hist.is_contiguous() and hist.dtype == x.dtype
    y = out if out is not None else torch.empty((N, Cout), dtype=x.dtype, device=x.device)
    _lib.call("wnb200_linear_step", _dt(x), N, Cin, Cout, k, int(dilation), int(step), _p(weight), _p(bias), _p(x),
              _p(hist), _p(y), _stream())
    return y


def channel_reduce(a, b=None, out=None):
    _need_cuda(a)
    a = a.contiguous()
    B, C, T = a.shape
    if out is None:
        out = torch.zeros(C, dtype=torch.float32, device=a.device)
    if b is not None:
        b = b.contiguous()
    _lib.call("wnb200_channel_reduce", _dt(a), B, C, T, _p(a), _p(b), _p(out), _stream())
    return out


def gate_bwd(dact, th, sg):
    _need_cuda(dact)
    dact = dact.contiguous()
    B, C, T = dact.shape
    dab = torch.empty((B, 2 * C, T), dtype=dact.dtype, device=dact.device)
    _lib.call("wnb200_gate_bwd", _dt(dact), B, C, T, _p(dact), _p(th), _p(sg), _p(dab), _stream())
    return dab


def gate_fwd(a, b, want_parts=False):
    _need_cuda(a, b)
    check_device()
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    th = torch.empty_like(a) if want_parts else None
    sg = torch.empty_like(a) if want_parts else None
    _lib.call("wnb200_gate_fwd", _dt(a), a.numel(), _p(a), _p(b), _p(out), _p(th), _p(sg), _stream())
    return (out, th, sg) if want_parts else out


def layernorm_bwd_params(x, stats, dy):
    B, C, T = x.shape
    dgamma = torch.zeros(C, dtype=torch.float32, device=x.device)
    dbeta = torch.zeros(C, dtype=torch.float32, device=x.device)
    _lib.call("wnb200_layernorm_bwd_params", _dt(x), B, C, T, _p(x), _p(stats), _p(dy.contiguous()), _p(dgamma),
              _p(dbeta), _stream())
    return dgamma, dbeta


def leaky_bwd(dy, ref):
    _need_cuda(dy, ref)
    dy = dy.contiguous()
    ref = ref.contiguous()
    dx = torch.empty_like(dy)
    _lib.call("wnb200_leaky_bwd", _dt(dy), dy.numel(), _p(dy), _p(ref), _p(dx), _stream())
    return dx


def softmax_fwd(x, log_mode=False):
    _need_cuda(x)
    check_device()
    x = x.contiguous()
    B, C, T = x.shape
    y = torch.empty_like(x)
    _lib.call("wnb200_softmax_fwd", _dt(x), B, C, T, _p(x), _p(y), int(log_mode), _stream())
    return y


def softmax_bwd(y, dy, log_mode=False):
    dy = dy.contiguous()
    B, C, T = y.shape
    dx = torch.empty_like(y)
    _lib.call("wnb200_softmax_bwd", _dt(y), B, C, T, _p(y), _p(dy), _p(dx), int(log_mode), _stream())
    return dx


def avgpool_fwd(x, pool):
    _need_cuda(x)
    check_device()
    x = x.contiguous()
    B, C, T = x.shape
    y = torch.empty((B, C, T // pool), dtype=x.dtype, device=x.device)
    _lib.call("wnb200_avgpool_fwd", _dt(x), B, C, T, pool, _p(x), _p(y), _stream())
    return y


def avgpool_bwd(dy, T, pool):
    dy = dy.contiguous()
    B, C, _ = dy.shape
    dx = torch.empty((B, C, T), dtype=dy.dtype, device=dy.device)
    _lib.call("wnb200_avgpool_bwd", _dt(dy), B, C, T, pool, _p(dy), _p(dx), _stream())
    return dx


def layernorm_fwd(x, gamma, beta, eps):
    _need_cuda(x)
    check_device()
    x = x.contiguous()
    B, C, T = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((B, T, 2), dtype=torch.float32, device=x.device)
    _lib.call("wnb200_layernorm_fwd", _dt(x), B, C, T, _p(x), _p(gamma), _p(beta), float(eps), _p(y), _p(stats),
              _stream())
    return y, stats


def layernorm_bwd(x, gamma, stats, eps, dy):
    dy = dy.contiguous()
    B, C, T = x.shape
    dx = torch.empty_like(x)
    _lib.call("wnb200_layernorm_bwd", _dt(x), B, C, T, _p(x), _p(gamma), _p(stats), float(eps), _p(dy), _p(dx),
              _stream())
    return dx


def xent_fwd(logits, target):
    _need_cuda(logits, target)
    check_device()
    logits = logits.contiguous()
    target = target.contiguous()
    assert target.dtype == torch.int64
    B, C, T = logits.shape
    loss_bt = torch.empty((B, T), dtype=torch.float32, device=logits.device)
    lse = torch.empty((B, T), dtype=torch.float32, device=logits.device)
    _lib.call("wnb200_xent_fwd", _dt(logits), B, C, T, _p(logits), _p(target), _p(loss_bt), _p(lse), _stream())
    return loss_bt, lse


def xent_bwd(logits, target, lse, gscale):
    B, C, T = logits.shape
    d = torch.empty_like(logits)
    _lib.call("wnb200_xent_bwd", _dt(logits), B, C, T, _p(logits), _p(target), _p(lse), _p(gscale), _p(d),
              _stream())
    return d


def sum_f32(x):
    _need_cuda(x)
    x = x.contiguous()
    out = torch.empty((), dtype=torch.float32, device=x.device)
    scratch = torch.empty(1024, dtype=torch.float32, device=x.device)
    _lib.call("wnb200_sum_f32", x.numel(), _p(x), _p(out), _p(scratch), _stream())
    return out


def positions_add_(out, w, bias, t0=0):
    B, F, T = out.shape
    _lib.call("wnb200_positions_add", _dt(out), B, F, T, int(t0), _p(w), _p(bias), _p(out), _stream())
    return out


def argmax_channels(x):
    _need_cuda(x)
    check_device()
    x = x.contiguous()
    B, C, T = x.shape
    out = torch.empty((B, T), dtype=torch.int64, device=x.device)
    _lib.call("wnb200_argmax_channels", _dt(x), B, C, T, _p(x), _p(out), _stream())
    return out


def _bct_strides(x, layout):
    """(B, L, T, sb, sc, st) of a 3-D activation tensor given as "bct" (network output), "btc" or "tbc"."""
    ix = {"bct": (0, 1, 2), "btc": (0, 2, 1), "tbc": (1, 2, 0)}[layout]
    return tuple(x.shape[i] for i in ix) + tuple(x.stride(i) for i in ix)


def frame_argmax(x, layout="bct"):
    """Per-frame argmax over the classes -> int64 (B, T); any of the three layouts is read in place."""
    _need_cuda(x)
    check_device()
    B, L, T, sb, sc, st = _bct_strides(x, layout)
    out = torch.empty((B, T), dtype=torch.int64, device=x.device)
    _lib.call("wnb200_frame_argmax", _dt(x), B, L, T, _p(x), sb, sc, st, _p(out), _stream())
    return out


def ctc_greedy_decode(x, act_lengths=None, blank=0, layout="bct"):
    """Argmax -> collapse repeats -> drop blanks, on the device.  Returns (labels int32 (B, T) with the decoded
    sequence of read b in labels[b, :lengths[b]], lengths int32 (B,))."""
    _need_cuda(x)
    check_device()
    B, L, T, sb, sc, st = _bct_strides(x, layout)
    labels = torch.zeros((B, T), dtype=torch.int32, device=x.device)
    lengths = torch.zeros((B,), dtype=torch.int32, device=x.device)
    al = None
    if act_lengths is not None:
        al = torch.as_tensor(act_lengths, dtype=torch.int32).to(x.device).contiguous()
    _lib.call("wnb200_ctc_greedy_decode", _dt(x), B, L, T, _p(x), sb, sc, st, _p(al), int(blank), _p(labels),
              _p(lengths), _stream())
    return labels, lengths
