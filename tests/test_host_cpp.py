"""The C++ host above the C ABI (ntrace_b200/host_cpp): FW::Buffer / RayBuffer / Scene / CudaBVH / HLBVHBuilder /
CudaBVHTracer / RayGen / CameraControls / Environment / Renderer mirrors and the runBenchmark driver `ntrace_bench`.

CPU: it compiles with g++ -Wall -Wextra against include/ntrace_b200.h, its GPU-free logic (config grammar, camera
signature codec and matrices, Wavefront ingestion order) agrees with the Python mirror, and without a GPU it fails loudly.
GPU: the benchmark binary runs the reference's loop on a scene file and its dumped rays / results check out against the
oracle; a bvhcache file it writes is read back by the Python mirror (same stream format)."""
import json
import os
import subprocess

import numpy as np
import pytest

from ntrace_b200 import camera, mesh_io, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ntrace_b200", "host_cpp")

CONFIG = """# reference-style config (grammar of src/rt/Environment.cpp, keys of config.conf)
App {
    benchmark true
    frameWidth 640   # comment after a value
    frameHeight 480
    stats %(stats)s
}
Benchmark {
    scene %(scene)s
    camera "%(camera)s"
    kernel b200_persistent_speculative_while_while
    warmupRepeats 1
    measureRepeats 2
}
Renderer { dataStructure BVH
    builder HLBVH
    rayType primary;AO;diffuse
    samples 4
    sortRays false
}
Raygen { aoRadius 2.5 }
SubdivisionRayCaster { numPrimitives 8 }   # a section this path does not know: must parse
"""

OBJ = """# quads, negative indices, materials, a face before any usemtl
mtllib test.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
v 1 0 1.25
vt 0 0
vn 0 0 1
f 1 2 3 4
usemtl red
f 1/1/1 2/1/1 5/1/1
f -1 -2 -3
usemtl unknown_material
f 2 3 6
usemtl blue
f 3//1 4//1 5//1 6//1
usemtl red
f 4 1 5
"""
MTL = "newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0 1\n"


@pytest.fixture(scope="module")
def built():
    import __graft_entry__  # noqa: F401  (the library must exist before the host links against it)
    from ntrace_b200 import build as nb
    nb.build()
    subprocess.run(["make", "-C", HOST, "-B", "all"], check=True, capture_output=True)
    return HOST


def test_cpp_host_logic_matches_python_mirror(built, tmp_path):
    (tmp_path / "test.obj").write_text(OBJ)
    (tmp_path / "test.mtl").write_text(MTL)
    conf = tmp_path / "config.conf"
    conf.write_text(CONFIG % dict(stats="stats.log", scene="scene.obj", camera=camera.SIGNATURES["conference"]))
    out = subprocess.run([os.path.join(built, "host_selftest"), str(conf), str(tmp_path / "test.obj"), "-DRenderer.samples=16", "-DBenchmark.kernel=a;b"],
                         check=True, capture_output=True, text=True).stdout
    got = json.loads(out)
    env = got["env"]
    assert env["App.frameWidth"] == "640" and env["App.frameHeight"] == "480"
    assert env["Renderer.samples"] == "16" and env["Benchmark.kernel"] == "a;b"                 # -D overrides the file
    assert env["Renderer.rayType"] == "primary;AO;diffuse" and env["Renderer.sortRays"] == "false"
    assert env["Benchmark.camera"].strip('"') == camera.SIGNATURES["conference"]
    assert env["SubdivisionRayCaster.numPrimitives"] == "8" and env["Raygen.aoRadius"] == "2.5"
    for name, c in got["cameras"].items():
        ref = camera.named_camera(name)
        assert np.array_equal(np.float32(c["position"]), ref.position) and np.array_equal(np.float32(c["forward"]), ref.forward)
        assert np.array_equal(np.float32(c["up"]), ref.up)
        assert (np.float32(c["fov"]), np.float32(c["near"]), np.float32(c["far"])) == (np.float32(ref.fov), np.float32(ref.near), np.float32(ref.far))
        m = np.float32(c["nscreenToWorld"]).reshape(4, 4)
        want = camera.nscreen_to_world(ref, 1024, 768)
        assert np.allclose(m, want, rtol=2e-5, atol=1e-6 * np.abs(want).max())
    verts, tris = mesh_io.load_obj(str(tmp_path / "test.obj"))
    assert np.array_equal(np.float32(got["verts"]), verts) and np.array_equal(np.int32(got["tris"]), tris)
    assert len(tris) == 8 and len(verts) == 13      # default submesh (3) + red (3) + blue (2); vertices split by (pos, tex, normal)


def test_cpp_host_rejects_bad_input(built, tmp_path):
    exe = os.path.join(built, "ntrace_bench")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Benchmark.scene is not set" in r.stderr
    bad = tmp_path / "bad.conf"
    bad.write_text("App { frameWidth 10 } }\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "unpaired }" in r.stderr
    r = subprocess.run([exe, "-DBenchmark.scene=x.obj", "-DBenchmark.camera=conference", "-DRenderer.dataStructure=KDTree"], capture_output=True, text=True)
    assert r.returncode == 1 and "Incorrect data structure type" in r.stderr


def test_cpp_host_has_no_cpu_fallback(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device error path cannot be shown")
    r = subprocess.run([os.path.join(built, "ntrace_bench"), "-DBenchmark.scene=x.obj", "-DBenchmark.camera=conference"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device available" in r.stderr
    # the host binary names no CUDA library itself: everything device-side goes through libntrace_b200.so
    needed = subprocess.run(["readelf", "-d", os.path.join(built, "ntrace_bench")], capture_output=True, text=True).stdout
    assert "libntrace_b200.so" in needed and "libcuda" not in needed.replace("libntrace_b200", "")


@pytest.mark.gpu
def test_cpp_benchmark_driver_end_to_end(built, orc, tmp_path):
    verts, tris = scenes.room(30_000, seed=13, wall_frac=0.3)
    mesh_io.save_ntmesh(str(tmp_path / "scene.ntmesh"), verts, tris)
    conf = tmp_path / "config.conf"
    conf.write_text(CONFIG % dict(stats=str(tmp_path / "stats.log"), scene=str(tmp_path / "scene.ntmesh"), camera=camera.SIGNATURES["conference"]))
    prefix = str(tmp_path / "dump")
    cache = str(tmp_path / "scene.bvhcache")
    r = subprocess.run([os.path.join(built, "ntrace_bench"), str(conf), f"-DBenchmark.dumpPrefix={prefix}", f"-DBenchmark.cacheFile={cache}",
                        "-DRenderer.cacheDataStructure=true"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Running benchmark for" in r.stdout and "[Mrays/s]" in r.stdout and "b200_persistent_speculative_while_while" in r.stdout
    table = [l for l in r.stdout.splitlines() if l.startswith("b200_persistent_speculative_while_while ")]
    mrays = [float(x) for x in table[-1].split()[1:]]
    assert len(mrays) == 3 and all(m > 50.0 for m in mrays)                        # a B200, not a CPU
    stats = open(tmp_path / "stats.log").read().split()
    assert stats.count("#SUM_RENDER_TIME") == 3 and stats.count("#SUM_RENDER_KRAYS") == 3
    krays = [float(stats[i + 1]) for i, s in enumerate(stats) if s == "#SUM_RENDER_KRAYS"]
    assert np.allclose(np.array(krays) * 1e-3, mrays, rtol=1e-2, atol=0.01)
    # dumped batches against the oracle (CPU SplitBVH of the same scene, traced on the dumped rays)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 8)
    flat = cpu.compact()
    k = "b200_persistent_speculative_while_while"
    cam = camera.named_camera("conference")
    prim = np.fromfile(f"{prefix}.{k}.primary.rays", dtype=np.float32).reshape(-1, 8)
    assert len(prim) == 640 * 480
    want_rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 640, 480), 640, 480, cam.far)
    assert np.allclose(prim, want_rays, rtol=1e-4, atol=1e-4)
    for rt, closest in (("primary", True), ("AO", False), ("diffuse", True)):
        rays = np.fromfile(f"{prefix}.{k}.{rt}.rays", dtype=np.float32).reshape(-1, 8)
        res = np.fromfile(f"{prefix}.{k}.{rt}.results", dtype=np.int32).reshape(-1, 4)
        assert len(rays) == len(res) > 0
        want = orc.compact_trace(*flat, rays, closest)                               # same Woop arithmetic as the kernel
        live = rays[:, 7] >= rays[:, 3]                                            # degenerate rays (missed primaries) carry no result
        if closest:
            same = (res[live, 0] == want[live, 0])
            assert same.mean() >= 0.9995
            hit = same & (want[live, 0] >= 0)
            tg, tw = res[live, 1].view(np.float32)[hit], want[live, 1].view(np.float32)[hit]
            # the driver's BVH is GPU-built: its Woop rows come from the builder's own transform (emitTreeKernel.cu:574-635
            # form), not from woopifyTri, so t agrees to rounding except on sliver triangles (numerics proper: test_gpu_build)
            rel = np.abs(tg - tw) / np.maximum(np.abs(tw), 1e-30)
            assert np.median(rel) <= 1e-6 and np.quantile(rel, 0.999) <= 5e-5
        else:
            assert ((res[live, 0] >= 0) == (want[live, 0] >= 0)).mean() >= 0.9995
    # the bvhcache file the C++ host wrote is the reference stream format: the Python mirror reads it and it is a valid tree
    from ntrace_b200 import host
    with open(cache, "rb") as f:
        bvh = host.CudaBVH.deserialize(f)
    assert bvh.layout == 4
    c = orc.canonical(bvh.nodes, bvh.woop, bvh.tri_index)
    assert sorted(c.tris.tolist()) == list(range(len(tris)))
    # second run loads the cache instead of building and gives the same primary results
    r2 = subprocess.run([os.path.join(built, "ntrace_bench"), str(conf), f"-DBenchmark.dumpPrefix={prefix}2", f"-DBenchmark.cacheFile={cache}",
                         "-DRenderer.cacheDataStructure=true", "-DRenderer.rayType=primary"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr
    a = np.fromfile(f"{prefix}.{k}.primary.results", dtype=np.int32).reshape(-1, 4)
    b = np.fromfile(f"{prefix}2.{k}.primary.results", dtype=np.int32).reshape(-1, 4)
    assert np.array_equal(a[:, :2], b[:, :2])


@pytest.mark.gpu
def test_named_cache_file_is_shared_by_both_hosts_and_pipelined_driver_runs(built, gpu_host, tmp_path):
    """Renderer::getCudaBVH's cache: "<cachePath>/<hash>_<builder>.dat" (Renderer.cpp:173-178).  The C++ driver writes it, the Python
    Renderer computes the same name for the same scene, imports the file instead of building, and traces the same results; the C++
    driver run a second time (pipelined loop) imports it too."""
    import re
    from ntrace_b200 import host
    verts, tris = scenes.room(12_000, seed=21, wall_frac=0.3)
    mesh = str(tmp_path / "scene.ntmesh")
    mesh_io.save_ntmesh(mesh, verts, tris)
    conf = tmp_path / "config.conf"
    conf.write_text(CONFIG % dict(stats=str(tmp_path / "stats.log"), scene=mesh, camera=camera.SIGNATURES["conference"]))
    cache_dir = str(tmp_path / "bvhcache")
    os.makedirs(cache_dir)
    args = [os.path.join(built, "ntrace_bench"), str(conf), f"-DBenchmark.cachePath={cache_dir}", "-DRenderer.cacheDataStructure=true"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    files = os.listdir(cache_dir)
    assert len(files) == 1 and re.fullmatch(r"[0-9a-f]{8}_HLBVH\.dat", files[0]), files
    scene = host.Scene(verts, tris)
    name = host.cache_file_name(scene, "HLBVH", host.BVHLayout_Compact, cache_dir)
    assert os.path.basename(name) == files[0]                      # one naming rule, two hosts
    # the Python Renderer imports the C++ host's file (no build: the handle has a host copy and no GPU build time) ...
    rnd = host.Renderer(host.BuildSettings(builder="HLBVH", cachePath=cache_dir))
    rnd.setScene(scene)
    bvh = rnd.getCudaBVH()
    assert not bvh.resident and bvh.gpu_seconds == 0.0 and bvh.nodes is not None
    # ... and traces what a freshly built tree of the same parameters traces
    cam = camera.named_camera("conference")
    rnd.setParams(host.RendererParams(rayType=host.RayType_Primary))
    rnd.beginFrame(cam, 320, 240)
    assert rnd.nextBatch()
    rnd.traceBatch()
    cached = rnd.m_batchRays.results_host().copy()
    fresh = host.Renderer(host.BuildSettings(builder="HLBVH"))
    fresh.setScene(scene)
    fresh.setParams(host.RendererParams(rayType=host.RayType_Primary))
    fresh.beginFrame(cam, 320, 240)
    assert fresh.nextBatch()
    fresh.traceBatch()
    assert np.array_equal(cached, fresh.m_batchRays.results_host())
    # second C++ run: imports the cache, pipelined batch loop
    mtime = os.path.getmtime(os.path.join(cache_dir, files[0]))
    r2 = subprocess.run(args + ["-DBenchmark.pipelined=true"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr
    assert os.listdir(cache_dir) == files and os.path.getmtime(os.path.join(cache_dir, files[0])) == mtime
    t1 = [l for l in r.stdout.splitlines() if l.startswith("b200_persistent_speculative_while_while ")][-1].split()[1:]
    t2 = [l for l in r2.stdout.splitlines() if l.startswith("b200_persistent_speculative_while_while ")][-1].split()[1:]
    assert len(t1) == len(t2) == 3 and all(float(x) > 50.0 for x in t2)


@pytest.mark.gpu
def test_cpp_driver_on_two_gpus_matches_one_gpu(built, tmp_path):
    """Renderer.numGpus in the C++ host: two processes, NCCL unique id through a file, rank 0 builds, nt_bvh_broadcast replicates, the
    secondary batches are dealt round-robin.  The frame's ray count is the single-GPU run's; throughput is reported by rank 0 only."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    verts, tris = scenes.room(30_000, seed=13, wall_frac=0.3)
    mesh = str(tmp_path / "scene.ntmesh")
    mesh_io.save_ntmesh(mesh, verts, tris)
    conf = tmp_path / "config.conf"
    conf.write_text(CONFIG % dict(stats=str(tmp_path / "stats1.log"), scene=mesh, camera=camera.SIGNATURES["conference"]))
    exe = os.path.join(built, "ntrace_bench")
    r1 = subprocess.run([exe, str(conf), "-DRenderer.samples=16", "-DBenchmark.pipelined=true"], capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0, r1.stderr
    comm = str(tmp_path / "nccl_id")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([exe, str(conf), "-DRenderer.samples=16", "-DBenchmark.pipelined=true", f"-DBenchmark.commFile={comm}",
                                       f"-DApp.stats={tmp_path / 'stats2.log'}"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    assert "2 GPUs: BVH broadcast" in outs[0][0] and "[Mrays/s]" in outs[0][0] and "[Mrays/s]" not in outs[1][0]

    def records(path):
        s = open(path).read().split()
        return [float(s[i + 1]) for i, x in enumerate(s) if x == "#SUM_RENDER_KRAYS"], [float(s[i + 1]) for i, x in enumerate(s) if x == "#SUM_RENDER_TIME"]
    k1, t1 = records(tmp_path / "stats1.log")
    k2, t2 = records(tmp_path / "stats2.log")
    assert len(k1) == len(k2) == 3
    # same frame, same ray accounting: rays = krays * time must agree between the runs for every ray type
    for a, b, c, d in zip(k1, t1, k2, t2):
        assert abs(a * b - c * d) <= 1e-3 * a * b
