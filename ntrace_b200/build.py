"""Build recipe for ``libntrace_b200.so`` (nvcc, sm_100a only, in-tree so it travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libntrace_b200.so")
SOURCES = ["nt_api.cu", "nt_trace.cu", "nt_raygen.cu", "nt_build.cu", "nt_raysort.cu", "nt_layout.cu", "nt_wide.cu", "nt_comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--threads", "0",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "ntrace_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def source_sha16() -> str:
    """sha256 (first 16 hex digits) over everything libntrace_b200.so is compiled from: csrc/*, the C header and the nvcc flags.  Ties an
    ncu capture under profiles/ to the code bench.py times (the .so itself is not bit-reproducible across nvcc runs)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    with open(os.path.join(_HERE, "..", "include", "ntrace_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart", "-ldl"]
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); let nvcc use the system host compiler
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB
