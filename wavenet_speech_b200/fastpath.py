"""Tensor-core (tcgen05 / TMA, NLC) inference pipeline behind the drop-in modules.

Two activation formats (include/wnb200.h, WNB200_ACT_*), chosen with `tc_precision("precise" | "fast")`:
  precise (default, C in {128, 256}): fp16 operands, the residual stream carried between layers as an fp16 (hi, lo) pair,
          exact gate -- holds the stated tolerance (2e-2 on the logits against the fp32 reference on the same
          bf16-rounded weights) at the depth of the benchmarked stacks (16-20 blocks);
  fast  : bf16 operands and stream, tanh.approx gate -- the format of the training path; ~2e-2 to about ten blocks.

A forward call is routed here when it is eligible: bf16 input, no autograd graph requested, every block
`C -> C` with C in {64, 128, 256} and kernel width <= 3.  Anything else continues on the generic NCL path
(`functional.py`), which is also the training path.  Weights are re-laid-out once per parameter version into
the K-major bf16 matrices the kernels stream by TMA (see wnb200.h, wnb200_chain_t):

  residual block l :  W1 = [Wtanh ; Wsigmoid] with tap-major columns          [2C, k*C]
                      W2 = [[Wres, Wproj], [Wbn_l @ Wskip, 0]]                 [2C, 2C]
                      b2 = [bres + bproj ; Wbn_l @ bskip + bbn_l]
  The skip -> bottleneck product is exact (reference wavenet.py:100 applies the bottleneck straight to
  conv1x1_skip's output) and saves one of the block's five contractions.
"""
import ctypes
import os

import torch

from . import _lib, ops

TC_GATE, TC_LEAKY, TC_LINEAR = 0, 1, 2
EPI2_RESBLOCK, EPI2_HEAD = 1, 2
_OK_C = (64, 128, 256)


def _pad16(n):
    return (n + 15) // 16 * 16


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f16(t):
    return t.detach().to(torch.float16).contiguous()


_PRECISE = True
_LOG2E = 1.4426950408889634


class tc_precision(object):
    """`tc_precision("precise")` / `tc_precision("fast")` as a statement, or as a context manager.
    Selects the activation format of the tensor-core INFERENCE path (see the module docstring)."""

    def __init__(self, mode="precise"):
        global _PRECISE
        assert mode in ("precise", "fast"), mode
        self.prev, _PRECISE = _PRECISE, mode == "precise"

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        global _PRECISE
        _PRECISE = self.prev
        return False


# ---- range guard of the precise format ---------------------------------------------------------------------------
# fp16 ends at 65504.  With the reference's initialisation the stream grows ~1.45x per block: 20 blocks reach ~1.5e3, but a
# 30-block 256-channel stack (the reference's tests/test_classifier.py shape) is at the limit and deeper ones beyond it.
# The block kernel stores saturated values (never inf) and raises a device flag; here the flag is read synchronously on
# the FIRST precise forward of every weight version (depth / initialisation problems show up there) -- if it is set the
# forward is repeated in the bf16 format and the model stays on it -- and asynchronously afterwards (a later,
# data-dependent overflow switches the model over from the next call on, with a warning).
_sat_ctx = None


def _sat_state(model, dev):
    st = model.__dict__.get("_wnb_sat")
    if st is None or st["flag"].device != dev:
        st = {"flag": torch.zeros(1, dtype=torch.int32, device=dev), "pinned": torch.zeros(1, dtype=torch.int32).pin_memory(),
              "event": None, "verified": None, "fast_key": None, "ran": False}
        model.__dict__["_wnb_sat"] = st
    return st


def _warn_saturated(model):
    import warnings
    warnings.warn("wavenet_speech_b200: the residual stream of %s left the fp16 range (|x| >= 65504) in the precise "
                  "tensor-core format; this model now runs in the bf16 format (tc_precision('fast') accuracy)"
                  % type(model).__name__, RuntimeWarning, stacklevel=3)


def precise_mode(C, model=None, dev=None):
    """The fp16 (hi, lo) format runs on the CTA-pair kernels only (C = 128 / 256).  With `model`: also consults / arms
    the range guard above (a model whose stream saturated fp16 stays on the bf16 format for that weight version)."""
    global _sat_ctx
    _sat_ctx = None
    ok = _PRECISE and C in (128, 256) and RESBLOCK_VARIANT != 1
    if not ok or model is None or dev is None:
        return ok
    key = _version_key(model)
    if torch.cuda.is_current_stream_capturing():
        # no allocations, event queries or read-backs inside a capture: a graph replays the format its warm-up chose
        # (GraphedForward warms up first); captured without a verified warm-up, the forward runs unguarded.
        st = model.__dict__.get("_wnb_sat")
        if st is not None and st["flag"].device == dev and st["fast_key"] == key:
            return False
        if st is not None and st["flag"].device == dev and st["verified"] == key:
            st["key"] = key
            _sat_ctx = st
        return True
    st = _sat_state(model, dev)
    if st["event"] is not None and st["event"].query():          # an earlier forward's asynchronous read-back
        st["event"] = None
        if int(st["pinned"][0]) != 0 and st["fast_key"] != key:
            st["fast_key"] = key
            _warn_saturated(model)
    if st["fast_key"] == key:
        return False
    if st["fast_key"] is not None or st["verified"] not in (None, key):     # new weights: start over
        st["fast_key"] = None
        st["verified"] = None
        st["flag"].zero_()
    st["key"] = key
    _sat_ctx = st
    return True


def stream_saturated(model):
    """True if the last checked forward of `model` (current weights) left the fp16 range and the model was moved to the
    bf16 format; False if its precise forward was verified in range; None if it has not run in the precise format."""
    st = model.__dict__.get("_wnb_sat")
    if st is None:
        return None
    key = _version_key(model)
    if st["fast_key"] == key:
        return True
    return False if st["verified"] == key else None


def range_guard(fn):
    """Wraps a try_*_forward: after a precise forward, act on the saturation flag (see above)."""
    import functools

    @functools.wraps(fn)
    def wrapper(model, x, *a, **k):
        global _sat_ctx
        y = fn(model, x, *a, **k)
        st, _sat_ctx = _sat_ctx, None
        if y is None or st is None or not st["ran"]:
            return y
        st["ran"] = False
        if torch.cuda.is_current_stream_capturing():
            return y
        if st["verified"] != st["key"]:
            if int(st["flag"].item()) != 0:                      # one synchronisation per weight version
                st["fast_key"] = st["key"]
                _warn_saturated(model)
                return fn(model, x, *a, **k)                     # precise_mode() now answers False for this model
            st["verified"] = st["key"]
        elif st["event"] is None:
            st["pinned"].copy_(st["flag"], non_blocking=True)
            st["event"] = torch.cuda.Event()
            st["event"].record()
        return y
    return wrapper


def _taps_matrix(w):
    """conv weight [M, C, k] -> [M, k*C] with tap-major columns."""
    M, C, k = w.shape
    return w.detach().float().permute(0, 2, 1).reshape(M, k * C)


def _sources(tensors):
    """The parameter tensors as the pack kernels read them: on the GPU, contiguous, ONE dtype (fp32 master weights or
    bf16; anything mixed is promoted to fp32)."""
    dt = tensors[0].dtype
    if dt not in (torch.float32, torch.bfloat16) or any(t.dtype != dt for t in tensors):
        dt = torch.float32
    out = []
    for t in tensors:
        t = t.detach()
        if not t.is_cuda:
            t = t.cuda()
        out.append(t.to(dt).contiguous())
    return out, ops._DT[dt]


def pack_block(block, bottleneck, precise=False, bwd=False, natural=True):
    """-> dict(w1, b1, w2, b2, offsets, ...) for one ResidualBlock + its skip bottleneck: ONE launch of
    `wnb200_pack_block` (the layouts are documented in wnb200.h), fold product Wbn * Wskip included.
    precise: fp16 weight matrices and gate biases pre-scaled for the ex2-based gate (WNB200_ACT_F16X2).
    bwd: also the transposed bf16 matrices of the two data-gradient contractions (training).
    natural: for C in {128, 256} also the [tanh ; sigmoid] row order of the generic chain kernel (keys w1 / b1) next to
    the pipelined kernels' order (w1h / b1h); the training step, which re-packs after every optimiser step, skips it."""
    C = block.out_channels
    wt, ws = block.conv_tanh.conv1d, block.conv_sigmoid.conv1d
    k = wt.weight.shape[2]
    src, wdt = _sources([wt.weight, wt.bias, ws.weight, ws.bias, block.conv1x1_residual.weight,
                         block.conv1x1_residual.bias, block.conv1x1_skip.weight, block.conv1x1_skip.bias,
                         block.residual_proj.weight, block.residual_proj.bias, bottleneck.weight, bottleneck.bias])
    dev = src[0].device
    wdtype = torch.float16 if precise else torch.bfloat16
    fmt = _lib.ACT_F16X2 if precise else _lib.ACT_BF16
    pk = {"offsets": list(block.offsets), "C": C, "fmt": fmt, "k": k}
    with torch.cuda.device(dev):
        orders = [(0, ("w1", "b1"))] if (natural and not precise) or C not in (128, 256) else []
        if C in (128, 256):
            orders.append((1, ("w1h", "b1h")))
        for order, names in orders:
            a = _lib.PackBlock()
            a.C, a.k, a.act_fmt, a.w_dtype, a.row_order = C, k, fmt, wdt, order
            for name, t in zip(("wt", "bt", "ws", "bs", "wres", "bres", "wskip", "bskip", "wproj", "bproj", "wbn", "bbn"),
                               src):
                setattr(a, name, t.data_ptr())
            w1 = torch.empty((2 * C, k * C), dtype=wdtype, device=dev)
            b1 = torch.empty(2 * C, dtype=torch.float32, device=dev)
            w2 = torch.empty((2 * C, 2 * C), dtype=wdtype, device=dev)
            b2 = torch.empty(2 * C, dtype=torch.float32, device=dev)
            a.w1, a.b1, a.w2, a.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
            if bwd and "wdg" not in pk:
                bf = torch.bfloat16
                pk["wdg"] = torch.empty((C, 2 * C), dtype=bf, device=dev)
                pk["wdg_skip"] = torch.empty((C, C), dtype=bf, device=dev)
                pk["wdx"] = torch.empty((C, k * 2 * C + C), dtype=bf, device=dev)
                pk["wdx_taps"] = torch.empty((C, k * 2 * C), dtype=bf, device=dev)
                a.wdg, a.wdg_skip, a.wdx, a.wdx_taps = (pk[n].data_ptr() for n in ("wdg", "wdg_skip", "wdx", "wdx_taps"))
            _lib.call("wnb200_pack_block", ctypes.byref(a), ops._stream())
            pk[names[0]], pk[names[1]], pk["w2"], pk["b2"] = w1, b1, w2, b2
    pk["_src"] = src            # the launches above read these asynchronously: keep them alive with the pack
    return pk


def pack_head(head, C, precise=False, bwd=False):
    """LeakyReLU -> 1x1 -> LeakyReLU -> 1x1: the first LeakyReLU is applied by the producer of the skip sum.  One launch
    of `wnb200_pack_head`; bwd adds the transposed bf16 matrices of the head's data gradients."""
    src, wdt = _sources([head[1].weight, head[1].bias, head[3].weight, head[3].bias])
    dev = src[0].device
    n_out = head[3].weight.shape[0]
    n2, npad = _pad16(n_out), (n_out + 63) // 64 * 64
    wdtype = torch.float16 if precise else torch.bfloat16
    fmt = _lib.ACT_F16X2 if precise else _lib.ACT_BF16
    pk = {"w1": torch.empty((C, C), dtype=wdtype, device=dev), "b1": torch.empty(C, dtype=torch.float32, device=dev),
          "w2": torch.empty((n2, C), dtype=wdtype, device=dev), "b2": torch.empty(n2, dtype=torch.float32, device=dev),
          "n_out": n_out, "n2": n2, "fmt": fmt, "npad": npad, "_src": src}
    a = _lib.PackHead()
    a.C, a.n_out, a.act_fmt, a.w_dtype = C, n_out, fmt, wdt
    a.w1, a.b1, a.w3, a.b3 = (t.data_ptr() for t in src)
    a.pw1, a.pb1, a.pw2, a.pb2 = (pk[n].data_ptr() for n in ("w1", "b1", "w2", "b2"))
    if bwd:
        pk["w3t"] = torch.empty((C, npad), dtype=torch.bfloat16, device=dev)
        pk["w1t"] = torch.empty((C, C), dtype=torch.bfloat16, device=dev)
        a.w3t, a.w1t = pk["w3t"].data_ptr(), pk["w1t"].data_ptr()
    with torch.cuda.device(dev):
        _lib.call("wnb200_pack_head", ctypes.byref(a), ops._stream())
    return pk


def _version_key(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def invalidate_packs(module):
    """Drop the cached weight packs of `module` and of every sub-module.  The cache is keyed on the parameters'
    (data_ptr, _version): in-place updates through the parameter (optimisers, load_state_dict, p.copy_) are seen, but
    writes through an alias with its own version counter -- `p.data.copy_(...)`, `p.data[...] = ...`, a raw-pointer write
    from outside torch -- are not.  Call this after such an update."""
    for m in module.modules():
        m.__dict__.pop("_wnb_pack_cache", None)
    from . import functional
    functional.clear_layout_cache()


def _cached(module, name, build):
    key = _version_key(module)
    cache = module.__dict__.setdefault("_wnb_pack_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        hit = (key, build())
        cache[name] = hit
    return hit[1]


def chain(x_nlc, C, offsets, epi1, w1, b1, n1, n2=0, use_x2=0, epi2=0, w2=None, b2=None, y_nlc=None, skips=None,
          skips_init=0, skips_act=None, out_ncl=None, n_out=0, softmax=0, dbg=None):
    _lib.current_tag = "resblock" if epi2 == EPI2_RESBLOCK else ("head" if epi2 == EPI2_HEAD else "dense")
    a = _lib.Chain()
    B, T, _ = x_nlc.shape
    a.B, a.T, a.C = B, T, C
    a.ntaps = len(offsets)
    for j, o in enumerate(offsets):
        a.t_off[j] = int(o)
    a.epi1, a.n1, a.n2, a.use_x2, a.epi2 = epi1, n1, n2, use_x2, epi2
    a.skips_init, a.n_out, a.softmax = skips_init, n_out, softmax
    a.out_f32 = 1 if (out_ncl is not None and out_ncl.dtype == torch.float32) else 0
    p = lambda t: 0 if t is None else t.data_ptr()
    a.x, a.w1, a.bias1, a.w2, a.bias2 = p(x_nlc), p(w1), p(b1), p(w2), p(b2)
    a.y_nlc, a.skips, a.skips_act, a.out_ncl = p(y_nlc), p(skips), p(skips_act), p(out_ncl)
    a.dbg = p(dbg)
    try:
        _lib.call("wnb200_chain_fwd_tc", ctypes.byref(a), ops._stream())
    finally:
        _lib.current_tag = None


RESBLOCK_VARIANT = int(__import__("os").environ.get("WNB200_RESBLOCK_VARIANT", "0"))


FUSE_FINAL = True     # last layer of an inference stack emits LeakyReLU(skip sum) as bf16 itself (no separate pass)


def resblock(x_nlc, pk, res, skips, skips_init, dbg=None, variant=None, save=None, skips_act=None, x_lo=None,
             res_lo=None, gate_out=None, sat_flag=None):
    """Pipelined fused block (C = 128 / 256).  variant 0/2 = CTA-pair kernel, 1 = single-CTA kernel.
    save = (gate, th, sg) NLC bf16 buffers: training keeps the gate and its two factors for backward.
    x_lo / res_lo: low halves of the fp16 (hi, lo) stream when the pack is a `precise` one."""
    a = _lib.ResBlock()
    B, T, C = x_nlc.shape
    a.B, a.T, a.C, a.ntaps = B, T, C, len(pk["offsets"])
    for j, o in enumerate(pk["offsets"]):
        a.t_off[j] = int(o)
    a.skips_init = int(skips_init)
    a.variant = RESBLOCK_VARIANT if variant is None else int(variant)
    p = lambda t: 0 if t is None else t.data_ptr()
    a.x, a.w1, a.bias1, a.w2, a.bias2 = p(x_nlc), p(pk["w1h"]), p(pk["b1h"]), p(pk["w2"]), p(pk["b2"])
    a.res, a.skips, a.dbg = p(res), p(skips), p(dbg)
    if save is not None:
        a.save_act, a.save_th, a.save_sg = p(save[0]), p(save[1]), p(save[2])
    a.skips_act = p(skips_act)
    a.act_fmt = pk.get("fmt", _lib.ACT_BF16)
    a.x_lo, a.res_lo = p(x_lo), p(res_lo)
    a.gate_out = p(gate_out)
    a.sat_flag = p(sat_flag)
    _lib.current_tag = "resblock"
    try:
        _lib.call("wnb200_resblock_fwd_tc", ctypes.byref(a), ops._stream())
    finally:
        _lib.current_tag = None


def dense(x_nlc, offsets, w, bias, N, mode=0, leaky=0, out=None, n_out=0, softmax=0, x2=None, offsets2=(), colsum=None,
          fmt=0, split=False, nlayers=0, gate_bwd=None, positions=None):
    """CTA-pair dense contraction.  mode 0 -> NLC bf16 [B,T,N]; mode 1 -> NCL `out` [B,n_out,T].
    Optional second source x2 [B,T,Cin2] with its own taps: its columns follow x's in `w`.
    colsum (mode 0): fp32 [N] tensor that the column sums of the output are ADDED to.
    nlayers: x_nlc is a stack [L, B, T, Cin] and w is [N, L*Cin]: y = epi(sum_l W_l x_l + bias), one contraction.
    gate_bwd = (gate, sigmoid) NLC bf16 [B,T,N]: the result is d(gate) and the epilogue applies the gate's backward:
    returns NLC bf16 [B,T,2N] = [d tanh-pre-activation | d sigmoid-pre-activation]; colsum is then fp32 [2N].
    positions = (w fp32 [N], b fp32 [N], t0): y += hardtanh(w * (t0 + t) + b) after the LeakyReLU (raw_ctcnet.py:131-135)."""
    if nlayers:
        assert x_nlc.dim() == 4 and x_nlc.shape[0] == nlayers and len(offsets) == 1 and x2 is None
        _L, B, T, Cin = x_nlc.shape
    else:
        B, T, Cin = x_nlc.shape
    a = _lib.Dense()
    a.B, a.T, a.Cin, a.ntaps = B, T, Cin, len(offsets)
    for j, o in enumerate(offsets):
        a.t_off[j] = int(o)
    a.N, a.mode, a.leaky, a.n_out, a.softmax = N, mode, int(leaky), n_out, int(softmax)
    out_lo = None
    if mode == 0:
        out = torch.empty((B, T, 2 * N if gate_bwd is not None else N),
                          dtype=torch.float16 if fmt == _lib.ACT_F16X2 else torch.bfloat16, device=x_nlc.device)
        if gate_bwd is not None:
            g, s = gate_bwd
            assert g.shape == (B, T, N) and s.shape == (B, T, N) and g.is_contiguous() and s.is_contiguous()
            assert g.dtype == torch.bfloat16 and s.dtype == torch.bfloat16
            a.gb_gate, a.gb_sg = g.data_ptr(), s.data_ptr()
        if split:
            assert fmt == _lib.ACT_F16X2
            out_lo = torch.empty_like(out)
            a.y_lo = out_lo.data_ptr()
    a.act_fmt = fmt
    if positions is not None:
        pw, pb, pt0 = positions
        assert pw.dtype == torch.float32 and pb.dtype == torch.float32 and pw.numel() == N and pb.numel() == N and mode == 0
        a.pos_w, a.pos_b, a.pos_t0 = pw.data_ptr(), pb.data_ptr(), int(pt0)
    a.nlayers = int(nlayers)
    a.out_f32 = 1 if out.dtype == torch.float32 else 0
    a.x, a.w, a.bias, a.y = x_nlc.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr()
    if colsum is not None:
        assert mode == 0 and colsum.dtype == torch.float32 and colsum.numel() >= (2 * N if gate_bwd is not None else N)
        a.colsum = colsum.data_ptr()
    if x2 is not None:
        assert x2.shape[:2] == x_nlc.shape[:2]
        a.x2, a.Cin2, a.ntaps2 = x2.data_ptr(), x2.shape[2], len(offsets2)
        for j, o in enumerate(offsets2):
            a.t_off2[j] = int(o)
    assert w.shape[1] == Cin * (nlayers or len(offsets)) + (0 if x2 is None else x2.shape[2] * len(offsets2)), \
        "dense: K mismatch"
    _lib.current_tag = "skipsum" if nlayers else ("dense" if mode == 0 else "head")
    try:
        _lib.call("wnb200_dense_fwd_tc", ctypes.byref(a), ops._stream())
    finally:
        _lib.current_tag = None
    return (out, out_lo) if split else out


def wgrad(g_nlc, x_nlc, off=0, m0=0, dw=None):
    """dw[m, n] += sum_{b,t} g[b,t,m0+m] * x[b,t+off,n] on tensor cores; returns dw fp32 [256, N]."""
    B, T, Cg = g_nlc.shape
    N = x_nlc.shape[2]
    if dw is None:
        dw = torch.zeros((256, N), dtype=torch.float32, device=g_nlc.device)
    _lib.current_tag = "wgrad"
    try:
        _lib.call("wnb200_wgrad_tc", B, T, Cg, int(m0), N, int(off), ops._p(g_nlc), ops._p(x_nlc), ops._p(dw),
                  ops._stream())
    finally:
        _lib.current_tag = None
    return dw


def wgrad2(g_nlc, xs, offs, m0=0, dw=None):
    """dw[m, s*N + n] = sum_{b,t} g[b,t,m0+m] * xs[s][b,t+offs[s],n] for one or two X tensors sharing the G tiles;
    returns fp32 [256, len(xs)*N]."""
    B, T, Cg = g_nlc.shape
    N = xs[0].shape[2]
    ns = len(xs)
    assert ns in (1, 2) and all(x.shape == xs[0].shape for x in xs)
    if dw is None:
        dw = torch.zeros((256, ns * N), dtype=torch.float32, device=g_nlc.device)
    off = (ctypes.c_int32 * 2)(int(offs[0]), int(offs[1]) if ns > 1 else 0)
    _lib.current_tag = "wgrad"
    try:
        _lib.call("wnb200_wgrad2_tc", B, T, Cg, int(m0), N, ns, off, ops._p(g_nlc), ops._p(xs[0]),
                  ops._p(xs[1] if ns > 1 else None), ops._p(dw), ops._stream())
    finally:
        _lib.current_tag = None
    return dw


WGRAD_MAXJOBS = 6


def wgrad_jobs(jobs):
    """jobs: list of (g_nlc, m0, xs, offs, dw) -- wgrad2's arguments -- over the SAME [B, T]: one launch for up to
    WGRAD_MAXJOBS of them (wnb200_wgrad_jobs_tc: the jobs sweep the frames together, shared operands come from L2)."""
    B, T = jobs[0][0].shape[:2]
    for i0 in range(0, len(jobs), WGRAD_MAXJOBS):
        part = jobs[i0:i0 + WGRAD_MAXJOBS]
        arr = (_lib.WgradJob * len(part))()
        for a, (g, m0, xs, offs, dw) in zip(arr, part):
            assert g.shape[:2] == (B, T) and len(xs) in (1, 2) and all(x.shape == xs[0].shape and x.shape[:2] == (B, T) for x in xs)
            a.g, a.Cg, a.m0 = g.data_ptr(), g.shape[2], int(m0)
            a.x, a.x2 = xs[0].data_ptr(), (xs[1].data_ptr() if len(xs) > 1 else None)
            a.N, a.nsrc = xs[0].shape[2], len(xs)
            a.off[0], a.off[1] = int(offs[0]), (int(offs[1]) if len(xs) > 1 else 0)
            assert dw.dtype == torch.float32 and dw.numel() >= 256 * len(xs) * xs[0].shape[2]
            a.dw = dw.data_ptr()
        _lib.current_tag = "wgrad"
        try:
            _lib.call("wnb200_wgrad_jobs_tc", B, T, len(part), ctypes.cast(arr, ctypes.c_void_p), ops._stream())
        finally:
            _lib.current_tag = None


def leaky_to_bf16(x):
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.call("wnb200_leaky_to_bf16", x.numel(), ops._p(x), ops._p(y), ops._stream())
    return y


def ncl_to_nlc_bf16(x, fmt=0):
    """(B, C, T) -> NLC bf16 (B, T, C) (fp16 with fmt = ACT_F16X2).  A tensor whose frames are contiguous (stride 1 along T, e.g. the time slice
    sig[:, :, 0:-1] of train.py:30) is read in place; anything else is made contiguous first."""
    B, C, T = x.shape
    if B * C * T > 0 and not (x.stride(2) == 1 and x.stride(1) >= T and x.stride(0) >= 0):
        x = x.contiguous()
    y = torch.empty((B, T, C), dtype=torch.float16 if fmt == _lib.ACT_F16X2 else torch.bfloat16, device=x.device)
    if B * C * T > 0:
        _lib.call("wnb200_ncl_to_nlc_act", ops._dt(x), fmt, B, C, T, x.stride(0), x.stride(1), ops._p(x), ops._p(y),
                  ops._stream())
    return y


def nlc_to_ncl(x, out_dtype):
    B, T, C = x.shape
    y = torch.empty((B, C, T), dtype=out_dtype, device=x.device)
    _lib.call("wnb200_nlc_to_ncl", ops._DT[out_dtype], 1 if x.dtype == torch.float32 else 0, B, C, T, ops._p(x),
              ops._p(y), ops._stream())
    return y


# fp32 models keep the fp32 FFMA kernels (<= 1e-5 against the reference) unless the caller opts in: with
# `reduced_precision(True)` fp32 inputs / parameters are routed through the tensor-core kernels too (bf16 operands, fp32
# accumulation, fp32 outputs and gradients; <= 2e-2 on logits like a bf16 model) -- the one-line switch for a reference
# checkpoint, which is fp32.
_REDUCED = False


class reduced_precision(object):
    """`reduced_precision(True)` as a statement, or `with reduced_precision(True): ...`."""

    def __init__(self, flag=True):
        global _REDUCED
        self.prev, _REDUCED = _REDUCED, bool(flag)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        global _REDUCED
        _REDUCED = self.prev
        return False


def tc_dtype_ok(t):
    return t.dtype == torch.bfloat16 or (_REDUCED and t.dtype == torch.float32)


def _no_graph(module, x):
    return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())))


def _stack_ok(C, layers):
    return C in _OK_C and all(ci == C and co == C and k <= 3 for (ci, co, k, _d) in layers)


DEFER_SKIP = True     # inference stacks (C = 128 / 256): gates stored per layer, the skip sum is ONE contraction at the end


def skip_pack(packs):
    """Operands of the stack-wide skip contraction: Wcat [C, L*C] = [fold_0 | fold_1 | ...] (fold_l = Wbn_l Wskip_l, the
    lower-left block of the layer's w2) and the summed skip biases.  Built once per weight version (cached by the caller)."""
    C = packs[0]["C"]
    wcat = torch.cat([pk["w2"][C:, :C] for pk in packs], 1).contiguous()
    bsum = torch.stack([pk["b2"][C:] for pk in packs]).sum(0).contiguous()
    return wcat, bsum


GATE_STACK_BUDGET = 48 << 30      # bytes; also capped at half of the free device memory


def _gate_stack_budget(dev):
    free, _total = torch.cuda.mem_get_info(dev)
    return min(GATE_STACK_BUDGET, free // 2)


# Precise format: how many of the LAST blocks of a stack carry the stream as its fp16 hi half only (the kernel accepts
# `res` without `res_lo`).  Dropping the lo half saves a seventh K slab and 2 x B*T*C*2 bytes of traffic per launch
# (0.065 ms of 0.47 per config-2 layer).  Measured at 20 WaveNet / 16 RawCTCNet blocks
# (profiles/r2_parity_hi_only_tail.jsonl, tests/tools/parity_hi_only_tail.py): worst output error 9.0e-3 with the pair
# everywhere, 1.1e-2 with 6 hi-only blocks, 1.3e-2 with 10, 1.6e-2 with all of them (forward 10.9 -> 10.6 -> 9.7 ms): the
# margin to the 2e-2 contract goes faster than the milliseconds, so the default keeps the pair in every block.
HI_ONLY_TAIL = int(os.environ.get("WNB200_HI_ONLY_TAIL", "0"))


def run_blocks_deferred(h, packs, skip, h_lo=None):
    """Residual stack with the skip sum deferred (resblock3_kernel + one `nlayers` contraction): returns the head's
    input LeakyReLU(skip sum) as an NLC tensor in the packs' format.  The running sum never exists in HBM.
    The gate stack costs L x B x T x C x 2 bytes (5.4 GB for config 2, 33.6 GB for the ecoli RawCTCNet at B = 1024): a
    batch whose stack would not fit the budget is walked in batch chunks (reads are independent)."""
    B, T, C = h.shape
    L = len(packs)
    need = L * B * T * C * 2
    budget = _gate_stack_budget(h.device)
    if need > budget and B > 1:
        nchunk = min(B, -(-need // max(budget, 1)))
        bounds = [(B * i // nchunk, B * (i + 1) // nchunk) for i in range(nchunk)]
        outs = [run_blocks_deferred(h[s:e], packs, skip, None if h_lo is None else h_lo[s:e]) for s, e in bounds if e > s]
        return torch.cat(outs, 0)
    prec = packs[0].get("fmt", 0) == _lib.ACT_F16X2
    gates = torch.empty((L, B, T, C), dtype=h.dtype, device=h.device)
    buf = [h, torch.empty_like(h)]
    lo = [h_lo, None]                              # [the input's lo half or None, a spare buffer]
    sat = None
    if prec and _sat_ctx is not None:
        sat = _sat_ctx["flag"]
        _sat_ctx["ran"] = True
    n_pair = max(0, L - 1 - HI_ONLY_TAIL)          # blocks 0 .. n_pair-1 hand a (hi, lo) pair to their successor
    for l, pk in enumerate(packs):
        last = l == L - 1
        out_lo = prec and not last and l < n_pair
        if out_lo and lo[1] is None:
            lo[1] = torch.empty_like(h)
        resblock(buf[0], pk, None if last else buf[1], None, False, x_lo=lo[0],
                 res_lo=lo[1] if out_lo else None, gate_out=gates[l], sat_flag=sat)
        if not last:
            buf = [buf[1], buf[0]]
            if prec:
                lo = [lo[1], lo[0]] if out_lo else [None, lo[1] if lo[1] is not None else lo[0]]
    wcat, bsum = skip
    return dense(gates, [0], wcat, bsum, C, leaky=1, fmt=packs[0].get("fmt", 0), nlayers=L)


def run_blocks(h, blocks, bottlenecks, packs, skips, first_init, want_act, h_lo=None, skip=None):
    """Residual stack on NLC activations: one fused launch per layer.  Returns (h, skips_act).
    Precise packs: `h` is the hi half of the fp16 stream, `h_lo` its lo half (None: the input is exactly `h`).
    skip = skip_pack(packs): use the deferred-skip pipeline (inference default for C = 128 / 256)."""
    B, T, C = h.shape
    if skip is not None and DEFER_SKIP and want_act and C in (128, 256) and RESBLOCK_VARIANT != 1 and (
            B > 1 or len(packs) * T * C * 2 <= _gate_stack_budget(h.device)):
        # (one read whose gate stack does not fit -- tens of millions of frames -- keeps the running sum in HBM instead)
        return None, run_blocks_deferred(h, packs, skip, h_lo=h_lo)
    if skips is None:
        skips = torch.empty((B, T, C), dtype=torch.float32, device=h.device)
    buf = [h, torch.empty_like(h)]
    n = len(packs)
    if C in (128, 256) and packs[0].get("fmt", 0) == _lib.ACT_F16X2:
        lo = [h_lo, torch.empty_like(h)]
        skips_act = torch.empty_like(h)
        for l, pk in enumerate(packs):
            last = l == n - 1
            resblock(buf[0], pk, None if last else buf[1], skips, first_init and l == 0,
                     skips_act=skips_act if last else None, x_lo=lo[0], res_lo=None if last else lo[1])
            if not last:
                buf = [buf[1], buf[0]]
                lo = [lo[1], lo[0] if lo[0] is not None else torch.empty_like(h)]
        return buf[0], skips_act
    if C in (128, 256):
        fuse = want_act and FUSE_FINAL and RESBLOCK_VARIANT != 1
        skips_act = torch.empty_like(h) if fuse else None
        for l, pk in enumerate(packs):
            last = l == n - 1
            resblock(buf[0], pk, None if last else buf[1], skips, first_init and l == 0,
                     skips_act=skips_act if last else None)
            if not last:
                buf = [buf[1], buf[0]]
        if fuse:
            return buf[0], skips_act
        return buf[0], (leaky_to_bf16(skips) if want_act else None)
    skips_act = torch.empty_like(h) if want_act else None
    for l, pk in enumerate(packs):
        last = l == n - 1
        chain(buf[0], C, pk["offsets"], TC_GATE, pk["w1"], pk["b1"], 2 * C, n2=2 * C, use_x2=1,
              epi2=EPI2_RESBLOCK, w2=pk["w2"], b2=pk["b2"], y_nlc=None if last else buf[1], skips=skips,
              skips_init=1 if (first_init and l == 0) else 0, skips_act=skips_act if last else None)
        if not last:
            buf = [buf[1], buf[0]]
    return buf[0], skips_act


def run_head(skips_act, hd, out_dtype, softmax):
    """LeakyReLU -> 1x1 -> LeakyReLU -> 1x1 [-> softmax] on the LeakyReLU'd skip sum: two dense launches."""
    B, T, C = skips_act.shape
    fmt = hd.get("fmt", 0)
    h1 = dense(skips_act, [0], hd["w1"], hd["b1"], C, leaky=1, fmt=fmt)
    out = torch.empty((B, hd["n_out"], T), dtype=out_dtype, device=skips_act.device)
    return dense(h1, [0], hd["w2"], hd["b2"], hd["n2"], mode=1, out=out, n_out=hd["n_out"], softmax=softmax, fmt=fmt)


@range_guard
def try_wavenet_forward(model, signal):
    """WaveNet.forward (reference wavenet.py:88-111) on the tensor-core path, or None if not eligible."""
    if not tc_dtype_ok(signal) or not signal.is_cuda or signal.dim() != 3:
        return None
    C = model.layers[0][0]
    if not _no_graph(model, signal):
        from . import training
        return training.wavenet_forward_train(model, signal) if training.wavenet_train_eligible(model, signal) else None
    if not (model.in_dim == C and model.out_dim == C and _stack_ok(C, model.layers)
            and model.entry_kwidth <= 3 and signal.shape[0] > 0 and signal.shape[2] > 0):
        return None
    ops.check_device()
    prec = precise_mode(C, model, signal.device)
    cast = _f16 if prec else _bf16
    fmt = _lib.ACT_F16X2 if prec else _lib.ACT_BF16

    def build():
        ec = model.entry_conv1d
        return {"entry": {"w1": cast(_taps_matrix(ec.conv1d.weight)), "b1": ec.conv1d.bias.detach().float().contiguous(),
                          "offsets": list(ec.offsets)},
                "blocks": [pack_block(b, n, prec) for b, n in zip(model.convolutions, model.bottlenecks)],
                "head": pack_head(model.output_stack, C, prec)}

    pk = _with_skip(_cached(model, "wavenet_p" if prec else "wavenet", build))
    B, _, T = signal.shape
    x = ncl_to_nlc_bf16(signal, fmt)
    h_lo = None
    if prec:
        h, h_lo = dense(x, pk["entry"]["offsets"], pk["entry"]["w1"], pk["entry"]["b1"], C, fmt=fmt, split=True)
    else:
        h = dense(x, pk["entry"]["offsets"], pk["entry"]["w1"], pk["entry"]["b1"], C)
    _, skips_act = run_blocks(h, model.convolutions, model.bottlenecks, pk["blocks"], None, True, True, h_lo=h_lo,
                              skip=pk["skip"])
    return run_head(skips_act, pk["head"], signal.dtype, model.softmax)


@range_guard
def wavenet_forward_levels(model, levels, out_dtype=torch.bfloat16):
    """WaveNet.forward(one_hot(levels)) without materialising the one-hot tensor: `levels` (B, T) integers in
    [0, in_dim) -- what pore_model.py:78-86 / np.digitize produce before the one-hot step (pore_model.py:88-96).  The entry
    conv becomes a gather of weight columns (`wnb200_entry_embed_nlc`), everything after it is the usual pipeline; the
    result is bit-identical to the one-hot call on the tensor-core path.  Inference only (no autograd graph)."""
    assert levels.is_cuda and levels.dim() == 2, "forward_levels expects a CUDA (B, T) integer tensor"
    C = model.layers[0][0]
    if not (model.in_dim == C and model.out_dim == C and C in (128, 256) and _stack_ok(C, model.layers)
            and model.entry_kwidth <= 3):
        raise RuntimeError("forward_levels: this WaveNet is not eligible for the tensor-core path "
                           "(in_dim = C = out_dim in {128, 256}, kernel widths <= 3)")
    ops.check_device()
    prec = precise_mode(C, model, levels.device)
    cast = _f16 if prec else _bf16
    fmt = _lib.ACT_F16X2 if prec else _lib.ACT_BF16

    def build():
        ec = model.entry_conv1d
        return {"wemb": cast(ec.conv1d.weight.detach().float().permute(2, 1, 0)),        # [k][in_dim][C]
                "b1": ec.conv1d.bias.detach().float().contiguous(), "offsets": list(ec.offsets),
                "blocks": [pack_block(b, n, prec) for b, n in zip(model.convolutions, model.bottlenecks)],
                "head": pack_head(model.output_stack, C, prec)}

    pk = _with_skip(_cached(model, "wavenet_levels_p" if prec else "wavenet_levels", build))
    B, T = levels.shape
    lev = levels.to(torch.int32).contiguous()
    h = torch.empty((B, T, C), dtype=torch.float16 if prec else torch.bfloat16, device=levels.device)
    h_lo = torch.empty_like(h) if prec else None
    if B * T > 0:
        offs = (ctypes.c_int32 * len(pk["offsets"]))(*[int(o) for o in pk["offsets"]])
        _lib.call("wnb200_entry_embed_nlc", B, T, C, model.in_dim, len(pk["offsets"]), offs, ops._p(lev),
                  ops._p(pk["wemb"]), ops._p(pk["b1"]), fmt, ops._p(h), ops._p(h_lo), ops._stream())
    with torch.no_grad():
        _, skips_act = run_blocks(h, model.convolutions, model.bottlenecks, pk["blocks"], None, True, True, h_lo=h_lo,
                                  skip=pk["skip"])
        return run_head(skips_act, pk["head"], out_dtype, model.softmax)


def _with_skip(pk):
    """Adds the stack-wide skip operands (Wcat, summed biases) to a cached pack dict (C = 128 / 256 stacks only)."""
    if "skip" not in pk:
        blocks = pk["blocks"]
        pk["skip"] = skip_pack(blocks) if blocks and blocks[0]["C"] in (128, 256) else None
    return pk


def _stack_packs(model, precise=False):
    packs = [pack_block(model.input_block, model.input_skip_bottleneck, precise)]
    packs += [pack_block(b, n, precise) for b, n in zip(model.convolutions, model.bottlenecks)]
    return packs


@range_guard
def try_raw_ctcnet_forward(model, seq, t0=0):
    """RawCTCNet.forward (reference raw_ctcnet.py:117-153) on the tensor-core path, or None if not eligible."""
    if not tc_dtype_ok(seq) or not seq.is_cuda or seq.dim() != 3 or seq.shape[1] != 1:
        return None
    C, F = model.layers[0][0], model.num_features
    if not _no_graph(model, seq):
        from . import training
        return training.raw_ctcnet_forward_train(model, seq) if training.raw_ctcnet_train_eligible(model, seq) else None
    if not (F == C and model.out_dim == C and C in (128, 256)
            and _stack_ok(C, model.layers) and model.input_kernel_size <= 3 and model.feature_kwidth * F <= 8192
            and seq.shape[0] > 0 and seq.shape[2] > 0):
        return None
    ops.check_device()
    prec = precise_mode(C, model, seq.device)
    cast = _f16 if prec else _bf16
    fmt = _lib.ACT_F16X2 if prec else _lib.ACT_BF16

    def build():
        f0, f2 = model.feature_layer[0], model.feature_layer[2]
        pos = None
        if model.positions:                       # position mixing rides in the epilogue of the feature layer's 1x1
            pc = model.positions_conv1x1[0]
            pos = (pc.weight.detach().float().reshape(-1).contiguous(), pc.bias.detach().float().contiguous())
        return {"f0w": f0.weight.detach().float()[:, 0, :].contiguous(), "f0b": f0.bias.detach().float().contiguous(),
                "f2w": cast(f2.weight.detach().float()[:, :, 0]), "f2b": f2.bias.detach().float().contiguous(), "pos": pos,
                "blocks": _stack_packs(model, prec), "head": pack_head(model.output_block, C, prec)}

    pk = _with_skip(_cached(model, "raw_ctcnet_p" if prec else "raw_ctcnet", build))
    B, _, T = seq.shape
    fk = model.feature_kwidth
    To = T + fk - 1
    seq = seq.contiguous()
    h = torch.empty((B, To, F), dtype=torch.float16 if prec else torch.bfloat16, device=seq.device)
    _lib.call("wnb200_featurize_nlc", ops._dt(seq), B, T, F, fk, ops._p(seq), ops._p(pk["f0w"]), ops._p(pk["f0b"]),
              fmt, ops._p(h), ops._stream())
    h_lo = None
    pos = None if pk["pos"] is None else (pk["pos"][0], pk["pos"][1], int(t0))
    if prec:
        h, h_lo = dense(h, [0], pk["f2w"], pk["f2b"], F, leaky=1, fmt=fmt, split=True, positions=pos)
    else:
        h = dense(h, [0], pk["f2w"], pk["f2b"], F, leaky=1, positions=pos)
    _, skips_act = run_blocks(h, None, None, pk["blocks"], None, True, True, h_lo=h_lo, skip=pk["skip"])
    return run_head(skips_act, pk["head"], seq.dtype, model.softmax)


@range_guard
def try_classifier_forward(model, seq):
    """WaveNetClassifier.forward (reference classifier.py:91-120) on the tensor-core path, or None."""
    if not tc_dtype_ok(seq) or not seq.is_cuda or seq.dim() != 3:
        return None
    C = model.layers[0][0]
    pool = model.pool_kernel_size
    if not _no_graph(model, seq):
        from . import training
        return training.classifier_forward_train(model, seq) if training.classifier_train_eligible(model, seq) else None
    if not (model.in_dim == C and model.out_dim == C and C in (128, 256)
            and _stack_ok(C, model.layers) and model.input_kernel_size <= 3 and seq.shape[0] > 0
            and seq.shape[2] // pool > 0):
        return None
    ops.check_device()
    prec = precise_mode(C, model, seq.device)
    fmt = _lib.ACT_F16X2 if prec else _lib.ACT_BF16
    pk = _with_skip(_cached(model, "classifier_p" if prec else "classifier",
                            lambda: {"blocks": _stack_packs(model, prec), "head": pack_head(model.output_block, C, prec)}))
    seq = seq.contiguous()
    B, _, T = seq.shape
    To = T // pool
    h = torch.empty((B, To, C), dtype=torch.float16 if prec else torch.bfloat16, device=seq.device)
    _lib.call("wnb200_avgpool_ncl_to_nlc", ops._dt(seq), B, C, T, pool, ops._p(seq), fmt, ops._p(h), ops._stream())
    _, skips_act = run_blocks(h, None, None, pk["blocks"], None, True, True, skip=pk["skip"])
    return run_head(skips_act, pk["head"], seq.dtype, model.softmax)


def smoke_check(reference_forward):
    """Tiny eligible WaveNet on the tensor-core path vs `reference_forward(state_dict, x, layers, softmax)`
    evaluated by the caller's checker on bf16-rounded weights (used by __graft_entry__.smoke())."""
    from .modules.wavenet import WaveNet
    torch.manual_seed(1)
    # 256 channels: the kernels the benchmark times (resblock2_kernel<256>, dense2_kernel), in both activation formats
    C = 256
    layers = [(C, C, 2, d) for d in (1, 2, 4, 8)]
    net = WaveNet(C, 2, layers, C, softmax=True)
    sd = {k: v.detach().bfloat16().float() for k, v in net.state_dict().items()}
    lev = torch.randint(0, C, (2, 300))
    x = torch.zeros(2, C, 300).scatter_(1, lev.unsqueeze(1), 1.0)
    ref = reference_forward(sd, x, layers, True)
    net = net.cuda().bfloat16()
    msgs = []
    for mode in ("precise", "fast"):
        with tc_precision(mode), torch.no_grad():
            y = try_wavenet_forward(net, x.cuda().bfloat16())
        assert y is not None, "tensor-core path refused an eligible shape"
        err = float((y.float().cpu() - ref).abs().max() / ref.abs().max())
        assert err < 2e-2, "tensor-core path (%s) mismatch vs the checker: %g" % (mode, err)
        msgs.append("%s %.2e" % (mode, err))
    return "tensor-core path (C=256) rel err " + ", ".join(msgs)
