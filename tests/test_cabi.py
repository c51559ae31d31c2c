"""The boundary is a C ABI: the header compiles as plain C (CPU check), and a C program linked against the library
builds a BVH and traces rays with no Python / torch in the process (GPU check)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_smoke.c")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _build(tmp_path):
    from ntrace_b200 import build
    build.build()
    exe = str(tmp_path / "cabi_smoke")
    pkg = os.path.join(ROOT, "ntrace_b200")
    subprocess.run([GCC, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", pkg, "-lntrace_b200", "-lm", f"-Wl,-rpath,{pkg}"], check=True, capture_output=True)
    return exe


def test_header_is_plain_c_and_program_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():
        # no GPU here: the program must fail loudly at nt_init (no CPU fallback), after passing the pre-init error check
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "nt_init" in r.stderr and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_program_builds_and_traces(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi_smoke OK" in r.stdout
