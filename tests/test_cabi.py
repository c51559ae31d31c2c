"""The boundary is a C ABI: the header compiles as plain C (CPU check), and a C program linked against the library
builds a BVH and traces rays with no Python / torch in the process (GPU check)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_smoke.c")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _build(tmp_path, src=SRC, name="cabi_smoke", std="-std=c99"):
    from ntrace_b200 import build
    build.build()
    exe = str(tmp_path / name)
    pkg = os.path.join(ROOT, "ntrace_b200")
    subprocess.run([GCC, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
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


MULTI = os.path.join(ROOT, "tests", "cabi", "cabi_multi.c")


def test_multi_gpu_c_program_links(tmp_path):
    assert os.path.exists(_build(tmp_path, MULTI, "cabi_multi", "-std=gnu99"))


def _run_multi(tmp_path, nranks):
    exe = _build(tmp_path, MULTI, "cabi_multi", "-std=gnu99")
    idfile = str(tmp_path / f"nccl_id_{nranks}")
    procs = [subprocess.Popen([exe, str(r), str(nranks), idfile], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(nranks)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    assert f"cabi_multi OK: {nranks} ranks" in outs[0][0]


@pytest.mark.gpu
def test_c_communicator_and_broadcast_single_rank(tmp_path):
    """NCCL bound by dlopen inside the library, communicator of one rank: the whole C path (unique id, init, broadcast, all-reduce)."""
    _run_multi(tmp_path, 1)


@pytest.mark.gpu
def test_c_two_processes_broadcast_and_shard(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    _run_multi(tmp_path, 2)
