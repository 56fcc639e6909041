"""ctypes binding of libwnb200.so (the C-ABI declared in include/wnb200.h).

The library is built in-tree by `wavenet_speech_b200/csrc/build.py` (nvcc, sm_100a).  There is no
CPU fallback: if the library cannot be loaded, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libwnb200_timeline.so" if os.environ.get("WNB200_TIMELINE", "0") == "1"
                        else "libwnb200.so")

c_void_p, c_int, c_int64, c_float_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p


ACT_BF16, ACT_F16X2 = 0, 1


class _Sized(ctypes.Structure):
    """Argument structs start with `struct_size` (= sizeof as this binding lays it out); the library refuses a mismatch."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.struct_size = ctypes.sizeof(self)


class Src(ctypes.Structure):
    """Mirror of wnb200_src_t."""
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("batch_stride", c_int64), ("chan_stride", c_int64),
                ("C", ctypes.c_int32), ("T_src", ctypes.c_int32), ("t_off", ctypes.c_int32),
                ("pre_act", ctypes.c_int32)]


class Ln(ctypes.Structure):
    """Mirror of wnb200_ln_t."""
    _fields_ = [("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p)]


class Taps(_Sized):
    """Mirror of wnb200_taps_t."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("dtype", ctypes.c_int32), ("B", ctypes.c_int32),
                ("T_out", ctypes.c_int32), ("M", ctypes.c_int32), ("nsrc", ctypes.c_int32),
                ("epilogue", ctypes.c_int32), ("accumulate", ctypes.c_int32), ("srcs", ctypes.POINTER(Src)),
                ("ln", ctypes.POINTER(Ln)), ("bias", c_void_p), ("out", c_void_p), ("th", c_void_p), ("sg", c_void_p),
                ("residual", c_void_p), ("mu_h", c_void_p), ("mu_h_ln", ctypes.c_int32), ("reserved0", ctypes.c_int32)]


class Chain(_Sized):
    """Mirror of wnb200_chain_t."""
    _fields_ = [("struct_size", ctypes.c_uint32),
                ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("C", ctypes.c_int32), ("ntaps", ctypes.c_int32),
                ("t_off", ctypes.c_int32 * 3), ("epi1", ctypes.c_int32), ("n1", ctypes.c_int32),
                ("n2", ctypes.c_int32), ("use_x2", ctypes.c_int32), ("epi2", ctypes.c_int32),
                ("skips_init", ctypes.c_int32), ("out_f32", ctypes.c_int32), ("n_out", ctypes.c_int32),
                ("softmax", ctypes.c_int32), ("x", c_void_p), ("w1", c_void_p), ("bias1", c_void_p),
                ("w2", c_void_p), ("bias2", c_void_p), ("y_nlc", c_void_p), ("skips", c_void_p),
                ("skips_act", c_void_p), ("out_ncl", c_void_p), ("dbg", c_void_p)]


class ResBlock(_Sized):
    """Mirror of wnb200_resblock_t."""
    _fields_ = [("struct_size", ctypes.c_uint32),
                ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("C", ctypes.c_int32), ("ntaps", ctypes.c_int32),
                ("t_off", ctypes.c_int32 * 3), ("skips_init", ctypes.c_int32), ("variant", ctypes.c_int32),
                ("act_fmt", ctypes.c_int32), ("x", c_void_p), ("w1", c_void_p),
                ("bias1", c_void_p), ("w2", c_void_p), ("bias2", c_void_p), ("res", c_void_p), ("skips", c_void_p),
                ("dbg", c_void_p), ("save_act", c_void_p), ("save_th", c_void_p), ("save_sg", c_void_p),
                ("skips_act", c_void_p), ("x_lo", c_void_p), ("res_lo", c_void_p), ("gate_out", c_void_p),
                ("sat_flag", c_void_p)]


class Dense(_Sized):
    """Mirror of wnb200_dense_t."""
    _fields_ = [("struct_size", ctypes.c_uint32),
                ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("Cin", ctypes.c_int32), ("ntaps", ctypes.c_int32),
                ("t_off", ctypes.c_int32 * 3), ("N", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("leaky", ctypes.c_int32), ("n_out", ctypes.c_int32), ("softmax", ctypes.c_int32),
                ("out_f32", ctypes.c_int32), ("Cin2", ctypes.c_int32), ("ntaps2", ctypes.c_int32),
                ("t_off2", ctypes.c_int32 * 3), ("x", c_void_p), ("w", c_void_p),
                ("bias", c_void_p), ("y", c_void_p), ("x2", c_void_p), ("colsum", c_void_p),
                ("act_fmt", ctypes.c_int32), ("nlayers", ctypes.c_int32), ("y_lo", c_void_p),
                ("gb_gate", c_void_p), ("gb_sg", c_void_p), ("pos_w", c_void_p), ("pos_b", c_void_p),
                ("pos_t0", ctypes.c_int32)]


class WgradJob(ctypes.Structure):
    """Mirror of wnb200_wgrad_job_t."""
    _fields_ = [("g", c_void_p), ("Cg", ctypes.c_int32), ("m0", ctypes.c_int32), ("x", c_void_p), ("x2", c_void_p),
                ("N", ctypes.c_int32), ("nsrc", ctypes.c_int32), ("off", ctypes.c_int32 * 2), ("dw", c_void_p)]


class AdamItem(ctypes.Structure):
    """Mirror of wnb200_adam_item_t."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64), ("is_bf16", ctypes.c_int32), ("reserved0", ctypes.c_int32)]


class PackBlock(_Sized):
    """Mirror of wnb200_pack_block_t."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("C", ctypes.c_int32), ("k", ctypes.c_int32),
                ("act_fmt", ctypes.c_int32), ("w_dtype", ctypes.c_int32), ("row_order", ctypes.c_int32)] + \
               [(n, c_void_p) for n in ("wt", "bt", "ws", "bs", "wres", "bres", "wskip", "bskip", "wproj", "bproj",
                                        "wbn", "bbn", "w1", "b1", "w2", "b2", "wdg", "wdg_skip", "wdx", "wdx_taps")]


class PackHead(_Sized):
    """Mirror of wnb200_pack_head_t."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("C", ctypes.c_int32), ("n_out", ctypes.c_int32),
                ("act_fmt", ctypes.c_int32), ("w_dtype", ctypes.c_int32), ("reserved0", ctypes.c_int32)] + \
               [(n, c_void_p) for n in ("w1", "b1", "w3", "b3", "pw1", "pb1", "pw2", "pb2", "w3t", "w1t")]


# name -> argtypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "wnb200_last_error": [],
    "wnb200_version": [],
    "wnb200_check_device": [],
    "wnb200_taps_fwd": [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(Src), c_void_p, c_int, c_int,
                        c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_taps_fwd_ex": [ctypes.POINTER(Taps), c_void_p],
    "wnb200_taps_wgrad_ex": [c_int, c_int, c_int, c_int, ctypes.POINTER(Src), ctypes.POINTER(Ln), c_void_p, c_void_p,
                             c_void_p],
    "wnb200_ln_stats": [c_int, c_int, c_int, c_int, c_void_p, ctypes.c_float, c_void_p, c_void_p],
    "wnb200_ln_relu_fwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_ln_relu_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_float, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_linear_frame": [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_linear_step": [c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p],
    "wnb200_lnrelu_rows": [c_int64, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float, c_void_p, c_void_p],
    "wnb200_mu_gate_rows": [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    "wnb200_nlc_parts_to_ncl_add": [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_void_p, c_void_p,
                                    c_void_p],
    "wnb200_adam_chunk_elems": [],
    "wnb200_adam_step": [c_int, c_void_p, c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                         ctypes.c_float, c_int64, c_void_p],
    "wnb200_taps_wgrad": [c_int, c_int, c_int, c_int, ctypes.POINTER(Src), c_void_p, c_void_p, c_void_p],
    "wnb200_channel_reduce": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_gate_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_gate_fwd": [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_layernorm_bwd_params": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p],
    "wnb200_leaky_bwd": [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_softmax_fwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "wnb200_softmax_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "wnb200_avgpool_fwd": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_avgpool_bwd": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_layernorm_fwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float,
                             c_void_p, c_void_p, c_void_p],
    "wnb200_layernorm_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float,
                             c_void_p, c_void_p, c_void_p],
    "wnb200_xent_fwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_xent_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_sum_f32": [c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_positions_add": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_positions_bwd": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_argmax_channels": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_chain_fwd_tc": [ctypes.POINTER(Chain), c_void_p],
    "wnb200_resblock_fwd_tc": [ctypes.POINTER(ResBlock), c_void_p],
    "wnb200_dense_fwd_tc": [ctypes.POINTER(Dense), c_void_p],
    "wnb200_pack_block": [ctypes.POINTER(PackBlock), c_void_p],
    "wnb200_pack_head": [ctypes.POINTER(PackHead), c_void_p],
    "wnb200_fold_grads": [c_int, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                          ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_workspace_bytes": [c_int, c_int, c_int, c_int],
    "wnb200_featurize_nlc": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "wnb200_avgpool_ncl_to_nlc": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "wnb200_avgpool_bwd_nlc_to_ncl": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_entry_embed_nlc": [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32), c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_wgrad_tc": [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_wgrad2_tc": [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p],
    "wnb200_wgrad_jobs_tc": [c_int, c_int, c_int, c_void_p, c_void_p],
    "wnb200_gate_bwd_nlc": [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_gate_bwd_nlc_from_gate": [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_colsum_nlc": [c_int64, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_ctc_workspace_bytes": [c_int, c_int, c_int, c_int],
    "wnb200_ctc_fwd": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_ctc_bwd": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "wnb200_frame_argmax": [c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p],
    "wnb200_ctc_greedy_decode": [c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int,
                                 c_void_p, c_void_p, c_void_p],
    "wnb200_featurize_bwd_nlc": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p],
    "wnb200_mu_gate_fwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_mu_gate_bwd": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_siggen_raw": [c_int, c_int, c_int, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_void_p],
    "wnb200_siggen_onehot": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "wnb200_leaky_to_bf16": [c_int64, c_void_p, c_void_p, c_void_p],
    "wnb200_ncl_to_nlc_bf16": [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "wnb200_ncl_to_nlc_bf16_strided": [c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p],
    "wnb200_ncl_to_nlc_act": [c_int, c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p],
    "wnb200_nlc_to_ncl": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
}
_RESTYPES = {"wnb200_last_error": ctypes.c_char_p, "wnb200_ctc_workspace_bytes": ctypes.c_size_t,
             "wnb200_workspace_bytes": ctypes.c_size_t}

_lib = None


class WnbError(RuntimeError):
    pass


def load():
    """Load (building first if the .so is absent and nvcc is available).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from .csrc import build as _build
        _build.build()
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise WnbError("wavenet_speech_b200: cannot load %s (%s); the CUDA library is required, there is "
                       "no CPU fallback" % (LIB_PATH, e))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise WnbError("wavenet_speech_b200: %s does not export %s (stale build?)" % (LIB_PATH, name))
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


# kernel launches issued per C-ABI call (for bench.py's `gpu_launches` claim)
_LAUNCHES_PER_CALL = {"wnb200_sum_f32": 2, "wnb200_last_error": 0, "wnb200_version": 0, "wnb200_check_device": 0,
                      "wnb200_tc_pack_bytes": 0, "wnb200_ctc_workspace_bytes": 0, "wnb200_ctc_fwd": 2,
                      "wnb200_workspace_bytes": 0, "wnb200_adam_chunk_elems": 0}
launch_count = 0
_event_log = None     # list of (name, start_event, end_event) while kernel timing is on
current_tag = None    # optional label (e.g. "resblock") attached to timed calls by the caller


def kernel_timing(enable):
    """Turn per-call CUDA-event timing on/off (events are recorded on torch's current stream, the stream the
    kernels are launched on).  Returns the log collected so far when turning off."""
    global _event_log
    log, _event_log = _event_log, ([] if enable else None)
    return log


def call(name, *args):
    global launch_count
    lib = load()
    fn = getattr(lib, name)
    if _event_log is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _event_log.append((name if current_tag is None else name + ":" + current_tag, e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        raise WnbError("%s failed (code %d): %s" % (name, rc, lib.wnb200_last_error().decode()))
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)
    return rc
