"""GPU: BVHs handed over in the reference's basic layouts (AOS_AOS, AOS_SOA, SOA_AOS, SOA_SOA; CudaBVH.cpp:453-575) are
rewritten on the device into the Compact form and must trace exactly like the Compact BVH of the same tree; the tesla_*
kernel names select AOS_AOS; malformed trees are refused; shadow-ray generation matches the oracle bit for bit."""
import os

import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu

LAYOUT_KERNELS = {0: "b200_persistent_speculative_while_while_aos_aos", 1: "b200_persistent_speculative_while_while_aos_soa",
                  2: "b200_persistent_speculative_while_while_soa_aos", 3: "b200_persistent_speculative_while_while_soa_soa"}
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def scene(orc):
    verts, tris = scenes.room(12_000, seed=9, wall_frac=0.3)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT, 1, 8)
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 256, 192), 256, 192, cam.far)
    rng = np.random.default_rng(5)
    lo, hi = scenes.bbox(verts)
    o = rng.uniform(lo, hi, (20000, 3)); d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rnd = np.concatenate([o, np.zeros((20000, 1)), d, np.full((20000, 1), 30.0)], axis=1).astype(np.float32)
    return verts, tris, cpu, np.ascontiguousarray(np.concatenate([rays, rnd]))


def _trace(host, bvh, kernel, rays, closest=True):
    tracer = host.CudaBVHTracer()
    tracer.setKernel(kernel)
    tracer.setBVH(bvh)
    rb = host.RayBuffer()
    rb.setRays(rays)
    rb.setNeedClosestHit(closest)
    tracer.traceBatch(rb)
    return rb.results_host()


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_basic_layouts_trace_like_compact(gpu_host, orc, scene, layout):
    verts, tris, cpu, rays = scene
    want = _trace(gpu_host, gpu_host.CudaBVH(*cpu.compact()), "b200_persistent_speculative_while_while", rays)
    nodes, woop, idx = cpu.basic(layout)
    got = _trace(gpu_host, gpu_host.CudaBVH(nodes, woop, idx, layout=layout), LAYOUT_KERNELS[layout], rays)
    assert np.array_equal(got[:, :2], want[:, :2])                      # same tree, same order: ids and t bit patterns
    assert (want[:, 0] >= 0).mean() > 0.5
    oracle_flat = orc.compact_trace(*cpu.compact(), rays, True)
    assert (got[:, 0] == oracle_flat[:, 0]).mean() >= 0.9999
    # any-hit agrees on hit / no hit
    any_got = _trace(gpu_host, gpu_host.CudaBVH(nodes, woop, idx, layout=layout), LAYOUT_KERNELS[layout], rays, closest=False)
    assert np.array_equal(any_got[:, 0] >= 0, want[:, 0] >= 0)
    # the caller's buffers are kept as given
    dn, dw, di, dl = capi.bvh_download()
    assert dl == layout and np.array_equal(dn, nodes) and np.array_equal(dw, woop) and np.array_equal(di, idx)


def test_convert_to_compact_is_the_same_tree(gpu_host, orc, scene):
    verts, tris, cpu, rays = scene
    ref_c = orc.canonical(*cpu.compact())
    for layout in (0, 3):
        nodes, woop, idx = cpu.basic(layout)
        capi.bvh_upload(layout, nodes, woop, idx)
        capi.bvh_convert(4)
        cn, cw, ci, cl = capi.bvh_download()
        assert cl == 4
        assert (cn.nbytes, cw.nbytes, ci.nbytes) == tuple(a.nbytes for a in cpu.compact())
        c = orc.canonical(cn, cw, ci)
        assert np.array_equal(c.inner, ref_c.inner) and np.array_equal(c.boxes.view(np.int32), ref_c.boxes.view(np.int32))
        assert np.array_equal(c.leaf_sizes, ref_c.leaf_sizes) and np.array_equal(c.tris, ref_c.tris)
        assert np.array_equal(c.woop.view(np.int32), ref_c.woop.view(np.int32))
        # Compact -> Compact2 -> Compact round trip of the links, and the kepler alias traces the Compact2 form
        capi.bvh_convert(5)
        n2 = capi.bvh_download()[0]
        link, link2 = cn.reshape(-1, 16)[:, 12:14], n2.reshape(-1, 16)[:, 12:14]
        assert np.array_equal(np.where(link >= 0, link // 16, link), link2)
        tracer = gpu_host.CudaBVHTracer()
        tracer.setKernel("kepler_dynamic_fetch")
        rb = gpu_host.RayBuffer(); rb.setRays(rays)
        bvh = gpu_host.CudaBVH(layout=5); bvh.resident = True
        tracer.setBVH(bvh)
        tracer.traceBatch(rb)
        want = orc.compact_trace(*cpu.compact(), rays, True)
        assert (rb.results_host()[:, 0] == want[:, 0]).mean() >= 0.9999
        capi.bvh_convert(4)
        assert np.array_equal(capi.bvh_download()[0], cn)


def test_tesla_kernel_names_select_aos_aos(gpu_host, orc, scene):
    verts, tris, cpu, rays = scene
    for name in ("tesla_persistent_while_while", "tesla_persistent_speculative_while_while", "tesla_persistent_packet"):
        tracer = gpu_host.CudaBVHTracer()
        tracer.setKernel(name)
        assert tracer.getDesiredBVHLayout() == 0
    got = _trace(gpu_host, gpu_host.CudaBVH(*cpu.basic(0), layout=0), "tesla_persistent_while_while", rays)
    want = orc.compact_trace(*cpu.compact(), rays, True)
    assert (got[:, 0] == want[:, 0]).mean() >= 0.9999
    # a Compact BVH under a tesla kernel is the reference's "Incorrect BVH layout!"
    tracer = gpu_host.CudaBVHTracer()
    tracer.setKernel("tesla_persistent_while_while")
    tracer.setBVH(gpu_host.CudaBVH(*cpu.compact()))
    rb = gpu_host.RayBuffer(); rb.setRays(rays[:64])
    with pytest.raises(gpu_host.NtError, match="Incorrect BVH layout"):
        tracer.traceBatch(rb)


def test_malformed_basic_bvh_is_refused(gpu_host, orc, scene):
    verts, tris, cpu, rays = scene
    nodes, woop, idx = cpu.basic(0)
    n = nodes.reshape(-1, 16)

    def upload(mut):
        m = nodes.copy()
        mut(m.reshape(-1, 16))
        capi.bvh_upload(0, m, woop, idx)

    with pytest.raises(capi.NtError, match="child index outside"):
        upload(lambda m: m.__setitem__((0, 12), len(m) + 5))
    with pytest.raises(capi.NtError, match="malformed BVH"):
        upload(lambda m: m.__setitem__((0, 12), 0))                                  # root as its own child
    inner_child = int(n[0, 12]) if n[0, 12] > 0 else int(n[0, 13])
    assert inner_child > 0
    with pytest.raises(capi.NtError, match="two parents"):
        upload(lambda m: (m.__setitem__((0, 12), inner_child), m.__setitem__((0, 13), inner_child)))
    leaf_rows = [~int(c) for c in n[:, 12:14].ravel() if c < 0]
    with pytest.raises(capi.NtError, match="leaf triangle range"):
        upload(lambda m: m.__setitem__((leaf_rows[0], 13), len(idx) + 1))
    with pytest.raises(capi.NtError, match="malformed BVH"):
        upload(lambda m: m.__setitem__((leaf_rows[1], 12), int(m[leaf_rows[0], 12])))  # two leaves start at the same triangle
    with pytest.raises(capi.NtError, match="inconsistent CudaBVH buffer sizes"):
        capi.bvh_upload(0, nodes, woop[:16], idx)
    with pytest.raises(capi.NtError, match="BVHLayout_CPU"):
        capi.bvh_upload(6, nodes, woop, idx)
    # a refused upload leaves no BVH behind
    rb = gpu_host.RayBuffer(); rb.setRays(rays[:32])
    with pytest.raises(capi.NtError, match="No BVH"):
        capi.trace_batch(rb.getRayBuffer(), rb.getResultBuffer(), 32, True)


def test_aos_upload_reproduces_frozen_reference_trace(gpu_host, orc):
    """AOS_AOS buffers (byte-identical to the reference's, tests/test_reference_pin.py) traced on the GPU against the
    frozen outputs of the reference's own CudaBVH::trace."""
    import json
    meta = json.load(open(os.path.join(HERE, "ref_golden.json")))
    d = np.load(os.path.join(HERE, "head.npz"))
    verts, tris = d["verts"], d["tris"]
    from test_reference_pin import pin_rays, sha
    rays = pin_rays(orc, verts)
    assert sha(rays) == meta["head"]["rays_sha"]
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 8, 1.0e-5)
    nodes, woop, idx = cpu.basic(0)
    assert [sha(nodes), sha(woop), sha(idx)] == meta["head"]["configs"]["sah_1_8"]["layout_sha"]["0"]
    got = _trace(gpu_host, gpu_host.CudaBVH(nodes, woop, idx, layout=0), "b200_persistent_speculative_while_while_aos_aos", rays)   # IEEE arithmetic
    want = np.load(os.path.join(HERE, "ref_head.npz"))["sah_1_8.flat"]
    same = got[:, 0] == want[:, 0]
    assert same.mean() >= 0.9999
    hit = same & (want[:, 0] >= 0)
    assert np.array_equal(got[hit, 1], want[hit, 1])


def test_raygen_shadow_matches_oracle(gpu_host, orc, scene):
    verts, tris, cpu, rays = scene
    prim = rays[: 256 * 192]
    res = orc.compact_trace(*cpu.compact(), prim, True)
    light, radius, seed, spp = (3.0, 12.0, 4.0), 0.75, 0xC0FFEE, 8
    irays = gpu_host.RayBuffer(); irays.setRays(prim)
    import torch
    irays.getResultBuffer().copy_(torch.from_numpy(res))
    gen = gpu_host.RayGen(maxBatchSize=100_000)
    out = gpu_host.RayBuffer()
    new_batch, done, total = True, 0, 0
    while True:
        ok, new_batch = gen.shadow(out, irays, spp, light, radius, new_batch, seed)
        if not ok:
            break
        n_in = out.getSize() // spp
        want, a, b = orc.raygen_shadow(prim, res, done, n_in, spp, light, radius, seed)
        got = out.rays_host()
        assert np.array_equal(got.view(np.int32), want.view(np.int32))
        assert np.array_equal(out.getIDToSlotBuffer().cpu().numpy(), a) and np.array_equal(out.getSlotToIDBuffer().cpu().numpy(), b)
        assert not out.getNeedClosestHit()
        done += n_in; total += out.getSize()
    assert done == len(prim) and total == len(prim) * spp
    # host-pointer path of the C ABI (staged copies) gives the same bytes
    host_out = np.zeros((1000 * spp, 8), np.float32); ia = np.zeros(1000 * spp, np.int32); ib = np.zeros(1000 * spp, np.int32)
    capi.raygen_shadow(host_out, ia, ib, prim, res, 500, 1000, spp, light, radius, seed)
    want, _, _ = orc.raygen_shadow(prim, res, 500, 1000, spp, light, radius, seed)
    assert np.array_equal(host_out.view(np.int32), want.view(np.int32))
