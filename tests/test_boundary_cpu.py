"""CPU-side checks of the drop-in boundary: state_dict compatibility with the reference (via the golden
fixtures, which hold reference state_dicts), constructor attributes, the C-ABI export list, and the
no-CPU-fallback rule."""
import ctypes
import os
import re

import pytest
import torch

import wavenet_speech_b200 as W
from wavenet_speech_b200 import _lib
from wavenet_speech_b200 import functional as WF
from tests import _golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wnb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wnb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.load()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the declared set
    assert set(_lib.SIGNATURES) == declared


def test_load_binds_and_reports_version():
    lib = _lib.load()
    assert lib.wnb200_version() >= 100


def test_state_dicts_load_reference_checkpoints_strictly():
    g = G.load("wavenet_onehot_c32")
    m = g["meta"]
    net = W.WaveNet(m["in_dim"], m["entry_kwidth"], m["layers"], m["out_dim"], softmax=m["softmax"])
    net.load_state_dict(g["sd"], strict=True)
    g = G.load("rawctcnet_positions")
    m = g["meta"]
    net = W.RawCTCNet(m["num_features"], m["feature_kwidth"], m["num_labels"], m["layers"], m["out_dim"],
                      positions=True, softmax=True)
    net.load_state_dict(g["sd"], strict=True)
    g = G.load("classifier_pool3")
    m = g["meta"]
    net = W.WaveNetClassifier(m["in_dim"], m["num_labels"], m["layers"], m["out_dim"],
                              pool_kernel_size=m["pool_kernel_size"], softmax=False)
    net.load_state_dict(g["sd"], strict=True)
    g = G.load("block_noncausal_k3_d3")
    blk = W.ResidualBlock(8, 8, 3, 3, causal=False)
    blk.load_state_dict(g["sd"], strict=True)
    g = G.load("layernorm_c6")
    W.LayerNorm(6).load_state_dict(g["sd"], strict=True)
    g = G.load("linearconv_k3_d2")
    W.LinearConv1d(4, 6, 3, dilation=2).load_state_dict(g["sd"], strict=True)
    g = G.load("multiplicative_unit")
    W.MultiplicativeUnit(6, 3, dilation=2).load_state_dict(g["sd"], strict=True)


def test_attributes_match_reference_semantics():
    c = W.CausalConv1d(4, 6, 5, dilation=3)
    assert c.padding == 12 and c.receptive_field == 13          # conv_ops.py:28,37
    n = W.NonCausalConv1d(4, 6, 2, dilation=3)
    assert n.padding == 2                                        # autopad(2,3) = ceil(3/2)
    assert W.autopad(3, 2) == 2 and W.autopad(2, 16) == 8 and W.autopad(2, 1) == 1
    b = W.ResidualBlock(4, 5, 2, 2, causal=False)
    assert b.receptive_field == 3 and b.causal is False and b.conditioning is False
    w = W.WaveNet(11, 2, [(11, 11, 2, d) for d in (1, 2, 4)], 11)
    assert (w.in_dim, w.out_dim, w.num_layers, w.softmax) == (11, 11, 3, True)
    lc = W.LinearConv1d(4, 6, 3, dilation=2)
    assert lc.receptive_field == 5 and lc._ker_ixs == [0, 2, 4]
    assert W.compute_new_length(15, 12, 3, 5) == 27.0


def test_tap_offsets():
    assert WF.tap_offsets(2, 1, False) == [-1, 0]
    assert WF.tap_offsets(2, 3, False) == [-2, 1]
    assert WF.tap_offsets(2, 16, False) == [-8, 8]
    assert WF.tap_offsets(2, 512, True) == [-512, 0]
    assert WF.tap_offsets(3, 2, False) == [-2, 0, 2]


def test_no_cpu_fallback():
    net = W.CausalConv1d(4, 6, 2, dilation=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.randn(1, 4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        W.LayerNorm(4)(torch.randn(1, 4, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wavenet_speech_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in _lib.py have the same size and field offsets as the C structs of include/wnb200.h
    (compiled here with gcc: the header is plain C and needs no CUDA)."""
    import ctypes
    import shutil
    import subprocess
    from wavenet_speech_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    pairs = {"wnb200_src_t": _lib.Src, "wnb200_chain_t": _lib.Chain, "wnb200_resblock_t": _lib.ResBlock,
             "wnb200_dense_t": _lib.Dense, "wnb200_ln_t": _lib.Ln, "wnb200_taps_t": _lib.Taps,
             "wnb200_pack_block_t": _lib.PackBlock, "wnb200_pack_head_t": _lib.PackHead,
             "wnb200_wgrad_job_t": _lib.WgradJob, "wnb200_adam_item_t": _lib.AdamItem}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "wnb200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ftype in cls._fields_:
            if fname.startswith("_"):
                continue                                    # explicit padding in the mirror
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        cname, fname, val = line.split()
        cls = pairs[cname]
        if fname == "sizeof":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), int(val))
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname, getattr(cls, fname).offset, int(val))


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument count and kind (pointer / 32-bit int / 64-bit int / float) of every ctypes binding against the header."""
    hdr = open(os.path.join(ROOT, "include", "wnb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char\s*\*)\s*(wnb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 15

    def c_kind(arg):
        arg = " ".join(arg.split())
        if "*" in arg:
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t|long long|size_t)\b", arg):
            return "i64"
        if re.search(r"\bfloat\b", arg):
            return "f32"
        return "i32"

    def py_kind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) not in (
                "i", "l", "q", "Q", "L", "I", "f"):
            return "ptr"
        return {"f": "f32"}.get(t._type_, "i64" if ctypes.sizeof(t) == 8 else "i32")

    seen = 0
    for name, args in protos:
        args = args.strip()
        kinds = [] if args in ("", "void") else [c_kind(a) for a in args.split(",")]
        bound = [py_kind(t) for t in _lib.SIGNATURES[name]]
        assert kinds == bound, (name, kinds, bound)
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_integration_doc_structs_match_the_binding():
    """INTEGRATION.md shows the ctypes structs a maintainer of another binding would mirror.  Round 1's copy had fallen
    behind the header (a missing trailing pointer): execute the documented class statements and compare them with
    `_lib.py` (itself checked against the header above), field by field."""
    import ctypes
    import re
    from wavenet_speech_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = text[text.index("```python\nimport ctypes, torch"):]
    block = block[len("```python\n"):block.index("\n```")]
    classes = re.findall(r"(class (wnb200_\w+_t)\(ctypes\.Structure\):.*?\n    _fields_ = \[.*?\]\n)\n", block, flags=re.S)
    assert {n for _, n in classes} == {"wnb200_pack_block_t", "wnb200_resblock_t"}
    ns = {"ctypes": ctypes, "vp": ctypes.c_void_p, "i32": ctypes.c_int32}
    for src, _name in classes:
        exec(src, ns)
    for name, mirror in (("wnb200_pack_block_t", _lib.PackBlock), ("wnb200_resblock_t", _lib.ResBlock)):
        doc = ns[name]
        assert [f[0] for f in doc._fields_] == [f[0] for f in mirror._fields_], name
        assert ctypes.sizeof(doc) == ctypes.sizeof(mirror), name
        for (n, t), (_n2, t2) in zip(doc._fields_, mirror._fields_):
            assert ctypes.sizeof(t) == ctypes.sizeof(t2), (name, n)
