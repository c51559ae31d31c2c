"""ntrace_b200 — B200-native (sm_100a) tracing / BVH-build path, drop-in for NTrace's tracing path.

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI (``include/ntrace_b200.h``),
``capi`` the ctypes binding, ``host`` the mirror of the reference's operator classes
(RayBuffer, CudaBVH, CudaBVHTracer, RayGen, Renderer), ``camera`` / ``scenes`` / ``environment``
host utilities.  There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401
from .capi import NtError  # noqa: F401

__all__ = ["capi", "NtError"]
