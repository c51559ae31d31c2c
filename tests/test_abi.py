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


def test_sass_listings_describe_the_library_as_built():
    """north_star: "each kernel is committed with its SASS listing".  profiles/sass/MANIFEST.json (scripts/dump_sass.sh) must name exactly
    the kernel symbols of the library built from the committed sources, every listed file must exist and start with that kernel's
    demangled name, and the manifest must have been generated from these sources."""
    import collections
    import json
    import re
    import shutil
    import subprocess
    from ntrace_b200 import build
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not available")
    lib_path = build.build()
    sass_dir = os.path.join(ROOT, "profiles", "sass")
    man = json.load(open(os.path.join(sass_dir, "MANIFEST.json")))
    assert man["source_sha16"] == build.source_sha16(), "kernel sources changed since profiles/sass was generated: run scripts/dump_sass.sh"
    txt = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, check=True).stdout
    mangled = re.findall(r"\n\s*Function : (\S+)", txt)
    dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")[: len(mangled)]
    have = collections.Counter()
    for d in dem:
        m = re.search(r"(\w+)<(.*)>\(", d) or re.search(r"(\w+)\(", d)
        targs = re.sub(r"\((?:int|bool|unsigned int)\)", "", m.group(2)) if m.lastindex and m.lastindex > 1 else ""
        have[m.group(1) + (f"<{targs}>" if targs else "")] += 1
    listed = collections.Counter(e["kernel"] for e in man["kernels"])
    assert have == listed
    for e in man["kernels"]:
        if e["listing"] is None:
            assert e.get("why")
            continue
        path = os.path.join(sass_dir, e["listing"])
        assert os.path.exists(path), path
        head = open(path).readline()
        assert e["kernel"].split("<")[0] in head
    # nothing stale lying around
    on_disk = {f for f in os.listdir(sass_dir) if f.endswith(".sass")}
    assert on_disk == {e["listing"] for e in man["kernels"] if e["listing"]}
