"""Host-side logic that needs no GPU: camera codec and matrices, pixel table, batching, bvhcache format,
slice partitioning, scene generators."""
import io

import numpy as np
import pytest

from ntrace_b200 import camera, multigpu, scenes


def test_camera_signatures_decode_to_survey_values():
    # SURVEY.md App. D (decoded with the codec of CameraControls.cpp:362-395,491-541)
    c = camera.named_camera("conference")
    assert np.allclose(c.position, [6.9742, 8.7685, 4.6396], atol=1e-4)
    assert np.allclose(c.forward, [0.7818, 0.5406, -0.3107], atol=1e-4)
    assert np.allclose(c.up, [0, 0, 1]) and abs(c.fov - 73.74) < 1e-2 and abs(c.near - 0.01) < 1e-6 and c.far == 100.0
    f = camera.named_camera("fairyforest")
    assert np.allclose(f.position, [0.0556, 0.2461, 0.5767], atol=1e-4) and abs(f.fov - 46.83) < 1e-2 and f.far == 500.0
    s = camera.named_camera("sibenik")
    assert np.allclose(s.position, [-19.1633, -1.5030, 4.6976], atol=1e-4)
    m = camera.named_camera("sanmiguel")
    assert np.allclose(m.forward, [0.0502, -0.0990, -0.9938], atol=1e-4) and abs(m.far - 126.0) < 0.01
    with pytest.raises(ValueError):
        camera.decode_signature("!!!")


def test_nscreen_to_world_inverts_the_projection(orc):
    cam = camera.named_camera("conference")
    w, h = 1024, 768
    n2w = camera.nscreen_to_world(cam, w, h)
    clip = (camera.fit_to_view(w, h) @ camera.perspective(cam.fov, cam.near, cam.far) @ camera.world_to_camera(cam)).astype(np.float32)
    assert np.allclose(n2w @ clip, np.eye(4), atol=2e-3)
    assert np.allclose(n2w, orc.invert4(clip), rtol=1e-5, atol=1e-4)       # same cofactor formula as Math.hpp:1024-1045
    assert abs(camera.fit_to_view(w, h)[0, 0] - 0.75) < 1e-7               # fov applies to the short side
    # centre pixel ray points along the camera forward axis
    rays, id2slot, slot2id = orc.raygen_primary(cam.position, n2w, 64, 48, cam.far)
    n2w_small = camera.nscreen_to_world(cam, 64, 48)
    rays, id2slot, slot2id = orc.raygen_primary(cam.position, n2w_small, 64, 48, cam.far)
    mid = rays[id2slot[24 * 64 + 32], 4:7]
    assert np.dot(mid, cam.forward / np.linalg.norm(cam.forward)) > 0.999


@pytest.mark.parametrize("w,h", [(1024, 768), (64, 48), (100, 75), (37, 21), (8, 8), (7, 5)])
def test_pixel_table_is_a_permutation_in_block_morton_order(orc, w, h):
    i2p, p2i = orc.pixel_table(w, h)
    assert sorted(i2p.tolist()) == list(range(w * h))
    assert np.array_equal(p2i[i2p], np.arange(w * h))
    if w >= 8 and h >= 8:
        first = i2p[:64]
        assert set((first % w).tolist()) == set(range(8)) and set((first // w).tolist()) == set(range(8))
        assert first[0] == 0 and first[1] == 1 and first[2] == w and first[3] == w + 1    # Morton inside the block


def test_batching_matches_reference_rule():
    from ntrace_b200.host import RayGen
    rg = RayGen(1 << 20)
    n, spp = 786_432, 32
    start, new, got = 0, True, []
    while True:
        ok, lo, hi, start, new = rg.batching(n, spp, start, new)
        if not ok:
            break
        got.append((lo, hi))
    assert got[0] == (0, 32768) and got[-1][1] == n and len(got) == 24
    assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
    assert rg.batching(0, spp, 0, True)[0] is False


def test_bvhcache_stream_format_roundtrip(orc):
    from ntrace_b200.host import CudaBVH
    verts, tris = scenes.soup_uniform(50, seed=2)
    nodes, woop, idx = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1).compact()
    b = CudaBVH(nodes, woop, idx, 4)
    buf = io.BytesIO()
    b.serialize(buf)
    raw = buf.getvalue()
    # S32 layout, then 3 x (S64 size, bytes), little endian (CudaBVH.cpp:105-125, Buffer.cpp:349-381)
    assert int.from_bytes(raw[:4], "little") == 4
    assert int.from_bytes(raw[4:12], "little") == nodes.nbytes
    assert len(raw) == 4 + 3 * 8 + nodes.nbytes + woop.nbytes + idx.nbytes
    c = CudaBVH.deserialize(io.BytesIO(raw))
    assert c.layout == 4 and np.array_equal(c.nodes, nodes) and np.array_equal(c.woop, woop) and np.array_equal(c.tri_index, idx)


@pytest.mark.parametrize("n,world", [(0, 4), (1, 4), (1_048_576, 8), (786_432, 3), (10, 16)])
def test_slices_partition_the_batch(n, world):
    parts = [multigpu.slice_for_rank(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert all(lo <= hi for lo, hi in parts)
    with pytest.raises(ValueError):
        multigpu.slice_for_rank(n, world, world)


def test_scene_generators_are_exact_and_deterministic():
    for gen, n in ((lambda: scenes.room(30_000, 2), 30_000), (lambda: scenes.teapot_in_stadium(20_000, 3), 20_000),
                   (lambda: scenes.soup_uniform(1_000, 5), 1_000)):
        v1, t1 = gen()
        v2, t2 = gen()
        assert len(t1) == n and t1.dtype == np.int32 and v1.dtype == np.float32
        assert np.array_equal(v1, v2) and np.array_equal(t1, t2)
        assert t1.min() >= 0 and t1.max() < len(v1)
    v, t, cam = scenes.config_scene("conference")
    assert len(t) == 283_000 and cam == "conference"
    lo, hi = scenes.bbox(v)
    c = camera.named_camera(cam)
    assert (c.position > lo).all() and (c.position < hi).all()      # the reference camera sits inside the stand-in room


def test_environment_reads_the_reference_config_grammar(tmp_path):
    from ntrace_b200.environment import Environment, EnvironmentError_
    text = """
App {
benchmark false
frameWidth 1024   # comment
}
Benchmark { scene data/Armadillo/armadillo.obj
camera GBSvz1V04qy/Ju69/21iChCz/idyKy10A0Kfx1pzUoy/DuY2/0aNqY10sZpuu/5/5f/0/
kernel fermi_speculative_while_while
warmupRepeats 1 }
Renderer {
# BVH construction currently supported = [SplitBVH, SAHBVH, OcclusionBVH, HLBVH, PersistentBVH]
builder PersistentKDTree
rayType primary
samples 8
sortRays true
}
SubdivisionRayCaster { numWarpsPerBlock 4
    Nested { depthK1 1.2 } }
"""
    p = tmp_path / "config.conf"
    p.write_text(text)
    env = Environment()
    env.Parse([str(p), "-DRenderer.builder=HLBVH", "-DRenderer.rayType=primary;AO", "-renderer_samples=32"])
    assert env.GetBool("App.benchmark") is False and env.GetInt("App.frameWidth") == 1024 and env.GetInt("App.frameHeight") == 768
    assert env.GetString("Renderer.builder") == "HLBVH" and env.GetInt("Renderer.samples") == 32
    assert env.GetString("Benchmark.kernel") == "fermi_speculative_while_while"
    assert env.GetString("SubdivisionRayCaster.Nested.depthK1") == "1.2"        # unknown sections are kept
    assert env.GetFloat("SBVH.alpha") == 1.0e-5 and env.GetBool("Renderer.sortRays") is True
    from ntrace_b200 import camera
    assert abs(camera.decode_signature(env.GetString("Benchmark.camera")).fov - 73.7) < 0.1
    with pytest.raises(EnvironmentError_):
        env.Parse(["-DApp.frameWidth=wide"])
    with pytest.raises(EnvironmentError_):
        Environment().ParseEnvString("App { benchmark true")
    with pytest.raises(EnvironmentError_):
        Environment().ParseEnvString("}")
    # the reference's shipped config.conf parses unchanged when present (not on the GPU box)
    import os
    if os.path.exists("/root/reference/config.conf"):
        e2 = Environment(); e2.ReadEnvFile("/root/reference/config.conf")
        assert e2.GetString("Renderer.dataStructure") == "KDTree" and e2.GetInt("Benchmark.measureRepeats") == 5


def test_gpu_numa_lookup_reads_sysfs(tmp_path):
    """multigpu.gpu_numa_node / _parse_cpulist: host placement of the pinned ray buffers next to the GPU."""
    from ntrace_b200 import multigpu
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    assert multigpu.gpu_numa_node(0, 0x1b, 0, sysfs=str(tmp_path)) == 1
    assert multigpu.gpu_numa_node(0, 0x2c, 0, sysfs=str(tmp_path)) == -1          # unknown device
    (dev / "numa_node").write_text("-1\n")
    assert multigpu.gpu_numa_node(0, 0x1b, 0, sysfs=str(tmp_path)) == -1          # platform does not say (VMs)
    assert multigpu._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert multigpu._parse_cpulist("") == set()


def test_shadow_ray_generator_properties(orc):
    """rayGenShadowKernel restated (RayGenKernels.cu:240-302): targets lie in the light's cube, tmax is the distance to
    the target, misses give degenerate rays, sample sets are Cranley-Patterson rotations of one QMC pattern."""
    rng = np.random.default_rng(3)
    n, spp = 200, 16
    o = rng.normal(size=(n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.zeros((n, 1)), d, np.full((n, 1), 100.0)], axis=1).astype(np.float32)
    res = np.zeros((n, 4), np.int32)
    t = rng.uniform(0.5, 5.0, n).astype(np.float32)
    res[:, 0] = np.where(np.arange(n) % 5 == 0, -1, np.arange(n))
    res[:, 1] = t.view(np.int32)
    light, radius = np.array([2.0, 9.0, -1.0], np.float32), 0.5
    out, a, b = orc.raygen_shadow(rays, res, 0, n, spp, light, radius, 77)
    assert np.array_equal(a, np.arange(n * spp)) and np.array_equal(b, np.arange(n * spp))
    out = out.reshape(n, spp, 8)
    origin = rays[:, :3] + rays[:, 4:7] * np.maximum(t - np.float32(1e-2), 0)[:, None]
    assert np.allclose(out[:, :, :3], origin[:, None, :], atol=1e-6)
    miss = res[:, 0] == -1
    assert (out[miss, :, 7] == -1.0).all() and (out[~miss, :, 7] > 0).all()
    assert np.allclose(np.linalg.norm(out[:, :, 4:7], axis=2), 1.0, atol=1e-5)
    target = out[~miss, :, :3] + out[~miss, :, 4:7] * out[~miss, :, 7:8]
    assert (np.abs(target - light) <= radius * (1 + 1e-4) + 1e-4).all()
    # per input ray the 16 targets are distinct and spread over the cube (z = Hammersley: one per 1/16 slab, rotated)
    z = np.sort(((target[:, :, 2] - light[2]) / radius + 1) / 2, axis=1)
    assert (np.diff(z, axis=1) > 0.03).all()
    # a different seed rotates the pattern
    out2, _, _ = orc.raygen_shadow(rays, res, 0, n, spp, light, radius, 78)
    assert not np.array_equal(out2.reshape(n, spp, 8)[~miss][:, :, 4:7], out[~miss][:, :, 4:7])


def test_bench_rank_ranges_cover_every_ray_exactly_once():
    """bench.py's multi-GPU split of a frame (both partitions): every slot of every ray type is traced by exactly one rank."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sizes = [("primary", 1000), ("AO", 640), ("AO", 640), ("AO", 640), ("AO", 77), ("diffuse", 640), ("diffuse", 13)]
    batches = [(t, np.arange(n * 8, dtype=np.float32).reshape(n, 8), n, t != "AO") for t, n in sizes]
    for partition in ("deal", "slices"):
        for world in (1, 2, 3, 4, 8):
            seen = {t: np.zeros(sum(n for tt, n in sizes if tt == t), dtype=np.int32) for t in ("primary", "AO", "diffuse")}
            for rank in range(world):
                for name, rays, n, closest, off in bench.rank_ranges(batches, rank, world, partition):
                    assert len(rays) == n and n > 0 and closest == (name != "AO")
                    seen[name][off:off + n] += 1
            for t in seen:
                assert (seen[t] == 1).all(), (partition, world, t)
