"""Build libpresight_b200.so (and the oracle-independent C-ABI) in-tree with nvcc for sm_100a.

    python -m presight_b200.build [--force]

The shared library lands in presight_b200/lib/ so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# debug variants (tools only): PS_LIB_SUFFIX=_dbg PS_NVCC_DEFS="-DPS_PHASE_CLOCKS" builds lib/libpresight_b200_dbg.so
SUFFIX = os.environ.get("PS_LIB_SUFFIX", "")
OBJDIR = os.path.join(HERE, "build" + SUFFIX)
LIB = os.path.join(LIBDIR, f"libpresight_b200{SUFFIX}.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
CFLAGS += os.environ.get("PS_NVCC_DEFS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "presight_b200.h"))
    return sorted(hs)


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def _compile(src, obj, log):
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs, hdrs = _sources(), _headers()
    stamp = os.path.join(OBJDIR, "stamp.txt")
    hdr_digest = _digest(hdrs)
    old = {}
    if os.path.exists(stamp) and not force:
        with open(stamp) as f:
            for line in f:
                k, v = line.strip().split(" ", 1)
                old[k] = v
    todo, objs, new = [], [], {}
    for s in srcs:
        name = os.path.basename(s)
        obj = os.path.join(OBJDIR, name.replace(".cu", ".o"))
        d = _digest([s]) + hdr_digest
        new[name] = d
        objs.append(obj)
        if force or old.get(name) != d or not os.path.exists(obj):
            todo.append((s, obj, os.path.join(OBJDIR, name + ".log")))
    if todo:
        if verbose:
            print(f"[presight_b200.build] compiling {len(todo)} file(s) for sm_100a ...", flush=True)
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda t: _compile(*t), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[presight_b200.build] linked {LIB}", flush=True)
    with open(stamp, "w") as f:
        for k, v in new.items():
            f.write(f"{k} {v}\n")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
