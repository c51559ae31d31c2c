"""Pins the restated oracle (oracle/liborc.so) against the REFERENCE ITSELF.

Two legs:
  * frozen:  tests/golden/ref_*.npz + ref_golden.json were produced by scripts/make_ref_golden.py from oracle/_ref/libref.so,
    i.e. the reference's own SAHBVHBuilder / SplitBVHBuilder / BVH::trace / CudaBVH (createCompact, woopifyTri, trace) /
    Intersect::* compiled unmodified from /root/reference.  The oracle must reproduce them bit for bit.  Runs anywhere.
  * live:    when /root/reference is present (build container) the same comparison runs on fresh seeded scenes.
The GPU leg (-m gpu) checks the CUDA traversal kernel against the frozen CudaBVH::trace outputs through the C ABI.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from ntrace_b200 import camera, scenes

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(HERE, "ref_golden.json")))
CONFIGS = {"sah_1_1": (False, 1, 1), "sah_1_8": (False, 1, 8), "split_1_1": (True, 1, 1), "split_1_8": (True, 1, 8)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def load(name):
    if name.startswith("room"):
        return scenes.room(5_000, seed=21)
    d = np.load(os.path.join(HERE, f"{name}.npz"))
    return d["verts"], d["tris"]


def fitted_camera(verts):
    lo, hi = scenes.bbox(verts)
    c = (lo + hi) * np.float32(0.5)
    d = float(np.linalg.norm(hi - lo))
    pos = c + np.array([0.45, 0.35, 0.3], np.float32) * np.float32(d)
    return camera.look_at(pos, c, up=(0.0, 1.0, 0.0), fov=60.0, near=d * 1e-3, far=d * 4.0)


def pin_rays(orc, verts, seed=7, n=4096):
    cam = fitted_camera(verts)
    prim, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    lo, hi = scenes.bbox(verts)
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    tmax = (rng.uniform(0.05, 1.5, (n, 1)) * np.linalg.norm(hi - lo)).astype(np.float32)
    rnd = np.concatenate([o, np.zeros((n, 1), np.float32), d, tmax], axis=1).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([prim, rnd], axis=0))


CASES = [(n, c) for n in ("map", "room5000_seed21", "head") for c in CONFIGS if not (n == "head" and c.startswith("split"))]


@pytest.mark.parametrize("name,cfg", CASES)
def test_oracle_reproduces_frozen_reference_outputs(orc, name, cfg):
    verts, tris = load(name)
    rays = pin_rays(orc, verts)
    assert sha(rays) == META[name]["rays_sha"], "pin rays drifted: regenerate with scripts/make_ref_golden.py"
    e = META[name]["configs"][cfg]
    split, mn, mx = CONFIGS[cfg]
    b = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if split else orc.BUILDER_SAH, mn, mx, 1.0e-5)
    st = b.stats()
    # builder: SAH cost (bit-exact double), node / leaf / reference counts, depth
    assert (st.sah, st.num_inner, st.num_leaf, st.num_tris, st.max_depth) == \
           (e["sah"], e["num_inner"], e["num_leaf"], e["num_tris"], e["max_depth"])
    # createCompact: Woop + index buffers byte-identical, tree identical up to the fork's node shuffle
    nodes, woop, idx = b.compact()
    assert sha(woop) == e["woop_buffer_sha"] and sha(idx) == e["tri_index_buffer_sha"]
    c = orc.canonical(nodes, woop, idx)
    assert (sha(c.inner), sha(c.boxes), sha(c.leaf_sizes)) == (e["inner_sha"], e["boxes_sha"], e["leaf_sizes_sha"])
    # the basic layouts (createNodeBasic / createTriWoopBasic / createTriIndexBasic): byte-identical buffers
    for L in range(4):
        assert [sha(a) for a in b.basic(L)] == e["layout_sha"][str(L)]
    # tracers: ids and t bit patterns, closest-hit and any-hit
    z = np.load(os.path.join(HERE, f"ref_{name}.npz"))
    assert np.array_equal(b.trace(rays, True)[:, :2], z[f"{cfg}.tree"])
    assert np.array_equal(b.trace(rays, False)[:, :2], z[f"{cfg}.tree_any"])
    assert np.array_equal(orc.compact_trace(nodes, woop, idx, rays, True)[:, :2], z[f"{cfg}.flat"])
    assert np.array_equal(orc.compact_trace(nodes, woop, idx, rays, False)[:, :2], z[f"{cfg}.flat_any"])
    assert int((z[f"{cfg}.flat"][:, 0] >= 0).sum()) == e["hits"] > 0


def test_oracle_primitives_match_frozen_reference(orc):
    z = np.load(os.path.join(HERE, "ref_primitives.npz"))
    rays, blo, bhi, tv = z["rays"], z["blo"], z["bhi"], z["tv"]
    box = np.stack([orc.ray_box(blo[i], bhi[i], rays[i]) for i in range(len(rays))])
    tri = np.stack([orc.ray_triangle(tv[i, 0:3], tv[i, 3:6], tv[i, 6:9], rays[i]) for i in range(len(rays))])
    inv = np.stack([orc.invert4(m) for m in z["mats"]])
    assert np.array_equal(box.view(np.int32), z["box_out"].view(np.int32))
    want = z["tri_out"]                                   # (t, u, v); the reference returns FLT_MAX in all three on a miss
    assert np.array_equal(tri[:, 0].view(np.int32), want[:, 0].view(np.int32))
    hit = want[:, 0] < np.float32(3.0e38)
    assert np.array_equal(tri[hit].view(np.int32), want[hit].view(np.int32))
    assert np.array_equal(inv.view(np.int32), z["inv_out"].view(np.int32))
    assert hit.sum() > 50                                 # the sample does contain hits


def test_oracle_pixel_table_matches_frozen_reference(orc):
    for key, want in META["pixel_table_sha"].items():
        w, h = map(int, key.split("x"))
        assert [sha(a) for a in orc.pixel_table(w, h)] == want


# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if r.build() is None:
        pytest.skip("no /root/reference here: live leg skipped, frozen leg above still pins the oracle")
    return r


@pytest.mark.parametrize("scene,seed", [("room", 3), ("soup", 4), ("teapot", 5)])
@pytest.mark.parametrize("cfg", ["sah_1_8", "split_1_8", "split_1_1"])
def test_oracle_matches_live_reference(orc, ref, scene, seed, cfg):
    verts, tris = {"room": lambda: scenes.room(2500, seed), "soup": lambda: scenes.soup_uniform(1500, seed),
                   "teapot": lambda: scenes.teapot_in_stadium(20000, seed)}[scene]()
    rays = pin_rays(orc, verts, seed=seed, n=2048)
    split, mn, mx = CONFIGS[cfg]
    r = ref.RefBVH(verts, tris, split=split, min_leaf=mn, max_leaf=mx)
    b = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if split else orc.BUILDER_SAH, mn, mx, 1.0e-5)
    rs, st = r.stats(), b.stats()
    assert (st.sah, st.num_inner, st.num_leaf, st.num_tris, st.max_depth) == \
           (rs["sah"], rs["num_inner"], rs["num_leaf"], rs["num_tris"], rs["max_depth"])
    rn, rw, ri = r.compact()
    nodes, woop, idx = b.compact()
    assert np.array_equal(rw, woop) and np.array_equal(ri, idx)
    rc, oc = orc.canonical(rn, rw, ri), orc.canonical(nodes, woop, idx)
    assert np.array_equal(rc.inner, oc.inner) and np.array_equal(rc.boxes.view(np.int32), oc.boxes.view(np.int32))
    assert np.array_equal(rc.leaf_sizes, oc.leaf_sizes) and np.array_equal(rc.tris, oc.tris)
    for L in range(4):
        for x, y in zip(r.layout(L), b.basic(L)):
            assert np.array_equal(x, y)
    assert np.array_equal(r.layout_trace(0, rays, True)[:, :2], r.compact_trace(rays, True)[:, :2])
    for closest in (True, False):
        assert np.array_equal(r.trace(rays, closest)[:, :2], b.trace(rays, closest)[:, :2])
        assert np.array_equal(r.compact_trace(rays, closest)[:, :2], orc.compact_trace(nodes, woop, idx, rays, closest)[:, :2])
    # the oracle's flat tracer also gives the reference's answers on the reference's own (shuffled) node buffer
    assert np.array_equal(r.compact_trace(rays, True)[:, :2], orc.compact_trace(rn, rw, ri, rays, True)[:, :2])
    assert np.array_equal(r.compact_trace(rays, True, nthreads=3), r.compact_trace(rays, True))


def test_live_reference_pixel_table(orc, ref):
    for w, h in ((640, 480), (129, 65), (16, 8), (9, 9)):
        a, b = ref.pixel_table(w, h), orc.pixel_table(w, h)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_live_reference_woop_primitive(orc, ref):
    rng = np.random.default_rng(11)
    for _ in range(500):
        v = rng.normal(size=(3, 3)).astype(np.float32)
        w = orc.woopify(v[0], v[1], v[2])
        ray = np.concatenate([rng.normal(size=3), [0.0], rng.normal(size=3), [100.0]]).astype(np.float32)
        got, want = orc.ray_triangle_woop(w, ray), ref.ray_triangle_woop(w, ray)
        assert got[0].view(np.int32) == want[0].view(np.int32)
        if want[0] < np.float32(3.0e38):
            assert np.array_equal(got.view(np.int32), want.view(np.int32))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg", [("map", "split_1_8"), ("room5000_seed21", "split_1_8"), ("head", "sah_1_8"), ("head", "sah_1_1")])
def test_gpu_kernel_reproduces_frozen_reference_trace(orc, gpu_host, name, cfg):
    """The CUDA traversal kernel, through nt_bvh_upload / nt_trace_batch, against CudaBVH::trace outputs of the reference."""
    verts, tris = load(name)
    rays = pin_rays(orc, verts)
    assert sha(rays) == META[name]["rays_sha"]
    split, mn, mx = CONFIGS[cfg]
    b = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if split else orc.BUILDER_SAH, mn, mx, 1.0e-5)
    nodes, woop, idx = b.compact()
    assert sha(woop) == META[name]["configs"][cfg]["woop_buffer_sha"]
    host = gpu_host
    bvh = host.CudaBVH(nodes, woop, idx)
    tracer = host.CudaBVHTracer()
    tracer.setBVH(bvh)
    rb = host.RayBuffer(len(rays))
    rb.setRays(rays)
    rb.setNeedClosestHit(True)
    tracer.traceBatch(rb)
    got = rb.results_host()[:, :2]
    want = np.load(os.path.join(HERE, f"ref_{name}.npz"))[f"{cfg}.flat"]
    same_id = got[:, 0] == want[:, 0]
    t_got, t_want = got[:, 1].view(np.float32), want[:, 1].view(np.float32)
    hit = want[:, 0] >= 0
    assert same_id.mean() >= 0.9999
    assert np.array_equal(t_got[hit & same_id].view(np.int32), t_want[hit & same_id].view(np.int32))   # t bit-exact
    bad = ~same_id
    assert np.all(np.abs(t_got[bad] - t_want[bad]) <= 1e-4 * np.maximum(1.0, np.abs(t_want[bad])))     # id differs only at t ties
    rb.setNeedClosestHit(False)
    tracer.traceBatch(rb)
    any_got = rb.results_host()[:, 0] >= 0
    assert np.array_equal(any_got, np.load(os.path.join(HERE, f"ref_{name}.npz"))[f"{cfg}.flat_any"][:, 0] >= 0)


# ---- bvhcache: the stream format and the file naming, pinned against the reference's own serializer / Hash.cpp -------------------------
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bvhcache_stream_written_by_the_reference_is_read_and_rewritten_byte_for_byte():
    """tests/golden/ref_map_compact.dat was written by the reference's CudaBVH::serialize (scripts/make_ref_cache_golden.py).  The Python
    reader must recover the three buffers and write the identical stream back; so must the C++ host's stream constructor."""
    import io
    import json
    import subprocess
    from ntrace_b200 import host
    data = open(os.path.join(GOLDEN_DIR, "ref_map_compact.dat"), "rb").read()
    kat = json.load(open(os.path.join(GOLDEN_DIR, "ref_cache_golden.json")))
    import hashlib
    assert hashlib.sha256(data).hexdigest() == kat["dat_sha256"]
    b = host.CudaBVH.deserialize(io.BytesIO(data))
    assert b.layout == 4 and b.nodes.nbytes % 64 == 0 and b.woop.nbytes == 4 * b.tri_index.nbytes
    assert 4 + 3 * 8 + b.nodes.nbytes + b.woop.nbytes + b.tri_index.nbytes == len(data)
    # it is the Map.obj SplitBVH: every one of the 488 triangles is referenced (triIndex holds [id, 0, 0] per triangle, 0 per terminator)
    assert set(b.tri_index.tolist()) == set(range(488))
    out = io.BytesIO()
    b.serialize(out)
    assert out.getvalue() == data
    exe = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "ntrace_b200", "host_cpp", "host_selftest")
    if os.path.exists(exe):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            dst = os.path.join(tmp, "out.dat")
            r = subprocess.run([exe, "--cache", os.path.join(GOLDEN_DIR, "ref_map_compact.dat"), dst], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            assert json.loads(r.stdout) == {"layout": 4, "nodeBytes": b.nodes.nbytes, "woopBytes": b.woop.nbytes, "idxBytes": b.tri_index.nbytes}
            assert open(dst, "rb").read() == data


def test_cache_file_name_hashes_match_the_reference_known_answers():
    """FW::hashBuffer (the library's nt_hash_buffer), hashBits and the cache-name formula against values computed by the reference's Hash.cpp."""
    import json
    import subprocess
    from ntrace_b200 import capi, host
    kat = json.load(open(os.path.join(GOLDEN_DIR, "ref_cache_golden.json")))
    data = open(os.path.join(GOLDEN_DIR, "ref_map_compact.dat"), "rb").read()
    for n, v in kat["hash_buffer"].items():
        assert capi.hash_buffer(data[: int(n)]) == v, n
    for n, v in kat["hash_buffer_unaligned"].items():
        assert capi.hash_buffer(data[1: 1 + int(n)]) == v, n
    assert host.hash_bits(*kat["scene_hash"]["args"]) == kat["scene_hash"]["value"]
    for row in kat["cache_name_hash"]:
        platform = host.hash_bits(capi.hash_buffer(b"GPU"), host._f2b(1.0), host._f2b(1.0), host.hash_bits(1, 1, row["minLeaf"], row["maxLeaf"]))
        params = host.hash_bits(host._f2b(row["splitAlpha"]))
        assert host.hash_bits(row["sceneHash"], platform, params, row["layout"], capi.hash_buffer(row["ds"].encode())) == row["value"]
    exe = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "ntrace_b200", "host_cpp", "host_selftest")
    if os.path.exists(exe):
        r = subprocess.run([exe, "--hash", os.path.join(GOLDEN_DIR, "ref_map_compact.dat")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = json.loads(r.stdout)
        assert got["hashBuffer"] == kat["hash_buffer"][str(len(data))]
        assert got == {"hashBuffer": got["hashBuffer"], "hashBits1": host.hash_bits(12345), "hashBits3": host.hash_bits(1, 2, 3),
                       "hashBits4": host.hash_bits(1, 2, 3, 4), "hashBits6": host.hash_bits(0xdeadbeef, 7, 0x80000000, 4, 5, 6)}


def test_live_reference_serializer_round_trip(ref):
    """With /root/reference present: a stream freshly written by the reference is read by the Python host, and the Python host's
    stream is accepted by the reference's own reader (CudaBVH(InputStream&), CudaBVH.cpp:105-108)."""
    import io
    from ntrace_b200 import host, scenes
    verts, tris = scenes.room(3000, seed=5)
    r = ref.RefBVH(verts, tris, split=False, min_leaf=1, max_leaf=4)
    for layout in (4, 5, 0):
        data = r.serialize(layout)
        n, w, i = r.layout(layout)
        b = host.CudaBVH.deserialize(io.BytesIO(data))
        assert b.layout == layout and np.array_equal(b.nodes, n) and np.array_equal(b.woop, w) and np.array_equal(b.tri_index, i)
        out = io.BytesIO()
        b.serialize(out)
        assert out.getvalue() == data
        L, nn, ww, ii = ref.deserialize(out.getvalue())
        assert L == layout and np.array_equal(nn, n) and np.array_equal(ww, w) and np.array_equal(ii, i)
