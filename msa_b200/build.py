"""Builds msa_b200/lib/libmmbert_sm100.so (sm_100a only) with nvcc.  No torch headers, no pybind: the
library exposes the plain C ABI of include/mmbert_sm100.h and is loaded with ctypes (msa_b200/capi.py)."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
OBJDIR = os.path.join(ROOT, "build")
LIB = os.path.join(LIBDIR, "libmmbert_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("MMB_NVCC_EXTRA", "").split()      # bring-up only, e.g. MMB_NVCC_EXTRA=-DMMB_ATTN_TRACE


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ["../../include/mmbert_sm100.h"]:
        fp = os.path.join(CSRC, f)
        if os.path.isfile(fp) and (f.endswith((".cuh", ".h")) or fp == path):
            h.update(open(fp, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return src, "cached", ""
    r = subprocess.run([NVCC] + FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return src, "built", r.stderr


def build(verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    rebuilt = any(s == "built" for _, s, _ in results)
    if verbose:
        for src, status, log in results:
            print(f"[{status}] {src}")
            if log:
                print(log)
    if rebuilt or not os.path.exists(LIB):
        objs = [os.path.join(OBJDIR, s[:-3] + ".o") for s in srcs]
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                           "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
