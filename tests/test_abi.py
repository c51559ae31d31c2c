"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ntrace_b200.h declares; calls fail loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ntrace_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nt_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from ntrace_b200 import build, capi
    build.build()
    l = capi.lib()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/ntrace_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu suite")
    import numpy as np
    from ntrace_b200 import NtError, capi
    with pytest.raises(NtError, match="no CUDA device|no CPU fallback|CUDA"):
        capi.init(0)
    rays = np.zeros((4, 8), np.float32)
    res = np.zeros((4, 4), np.int32)
    with pytest.raises(NtError, match="nt_init"):
        capi.trace_batch(rays, res, 4, True)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "ntrace_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liborc" not in text and "orc_" not in text, f
