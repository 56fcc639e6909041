"""Build libwnb200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a plain C-ABI)."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
# WNB200_TIMELINE=1: diagnostic build with the clock64 phase stamps compiled in (scripts/timeline.py), kept apart
# from the shipped library
TIMELINE = os.environ.get("WNB200_TIMELINE", "0") == "1"
LIB = os.path.join(PKG, "libwnb200_timeline.so" if TIMELINE else "libwnb200.so")
STAMP = os.path.join(PKG, ".libwnb200_timeline.stamp" if TIMELINE else ".libwnb200.stamp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + (["-DWNB200_TIMELINE"] if TIMELINE else [])


def _sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(HERE, "*.cuh"))) + [
            os.path.join(os.path.dirname(PKG), "include", "wnb200.h"), __file__]:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _file_digest(src, headers):
    h = hashlib.sha256()
    for p in [src] + headers + [__file__]:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """One `nvcc -c` per .cu (in parallel, objects cached under csrc/_obj by content hash), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objdir = os.path.join(HERE, "_obj_timeline" if TIMELINE else "_obj")
    os.makedirs(objdir, exist_ok=True)
    headers = sorted(glob.glob(os.path.join(HERE, "*.cuh"))) + [os.path.join(os.path.dirname(PKG), "include", "wnb200.h")]
    cflags = [f for f in FLAGS if f != "--shared"]
    logs = {}

    def compile_one(src):
        base = os.path.splitext(os.path.basename(src))[0]
        obj, stamp = os.path.join(objdir, base + ".o"), os.path.join(objdir, base + ".stamp")
        d = _file_digest(src, headers)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == d:
            return obj, 0
        cmd = [NVCC] + cflags + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs[base] = " ".join(cmd) + "\n" + res.stdout + res.stderr
        if res.returncode == 0:
            with open(stamp, "w") as f:
                f.write(d)
        return obj, res.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = list(ex.map(compile_one, _sources()))
    link = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + [o for o, _ in results]
    bad = [o for o, rc in results if rc != 0]
    res = None if bad else subprocess.run(link, capture_output=True, text=True)
    log = "".join(logs[k] for k in sorted(logs)) + ("" if res is None else " ".join(link) + "\n" + res.stdout + res.stderr)
    prev = os.path.join(PKG, "build.log")
    with open(prev, "a" if logs and os.path.exists(prev) and not force else "w") as f:
        f.write(log)
    if bad or res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libwnb200.so")
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
