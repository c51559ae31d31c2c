"""The restated LBVH emitter against an independent pure-Python statement of the range rule,
plus Morton / sort invariants and HLBVH sanity."""
import numpy as np
import pytest

from ntrace_b200 import scenes


def _py_morton(verts, tris, lo, hi):
    v = verts.astype(np.float32)
    a, b, c = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    tlo = np.minimum(a, np.minimum(b, c)); thi = np.maximum(a, np.maximum(b, c))
    mid = (tlo + (thi - tlo) / np.float32(2.0)).astype(np.float32)
    step = ((hi - lo) / np.float32(1024.0)).astype(np.float32)
    with np.errstate(all="ignore"):
        q = np.floor(((mid - lo) / step).astype(np.float32))
    q = np.nan_to_num(q, nan=0.0, posinf=2 ** 31, neginf=-2 ** 31)
    q = np.clip(q, 0, 1023).astype(np.uint32)

    def spread(n):
        n = n & 0x3ff
        n = (n ^ (n << 16)) & 0xff0000ff
        n = (n ^ (n << 8)) & 0x0300f00f
        n = (n ^ (n << 4)) & 0x030c30c3
        return (n ^ (n << 2)) & 0x09249249
    return (spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)).astype(np.uint32)


def _py_range_tree(keys, leaf):
    """(s, split, e, leftLeaf, rightLeaf, word14) per inner node in preorder, by the literal level rule."""
    out = []

    def rec(s, e, level):
        b0 = 29 - level
        b = b0
        while b >= 0 and ((int(keys[s]) >> b) & 1) == ((int(keys[e - 1]) >> b) & 1):
            b -= 1
        if b >= 0:
            first = (int(keys[s]) >> b) & 1
            split = next(i for i in range(s + 1, e) if ((int(keys[i]) >> b) & 1) != first)
        else:
            split = (s + e) >> 1
        ll = (split - s) <= leaf or b0 == 0
        rl = (e - split) <= leaf or b0 == 0
        w14 = -1 if b < 0 else b % 3
        slot = len(out)
        out.append(None)
        cl = (split - s) if ll else rec(s, split, level + 1)
        cr = (e - split) if rl else rec(split, e, level + 1)
        out[slot] = (cl, cr, w14)
        return cl + cr
    rec(0, len(keys), 0)
    return np.array(out, dtype=np.int32)


@pytest.mark.parametrize("n,leaf,seed", [(500, 8, 1), (500, 1, 2), (3000, 4, 3), (7, 8, 4), (2, 1, 5), (1, 8, 6)])
def test_emitter_matches_python_rule(orc, n, leaf, seed):
    verts, tris = scenes.soup_uniform(n, seed=seed, clustered=(seed % 2 == 1))
    lo, hi = scenes.bbox(verts)
    lo, hi = lo - np.float32(8.0), hi + np.float32(8.0)            # coarse grid -> duplicate codes
    codes = orc.morton(verts, tris, lo, hi)
    assert np.array_equal(codes, _py_morton(verts, tris, lo, hi))
    r = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf)
    order = np.argsort(codes, kind="stable")
    assert np.array_equal(r.sorted_idx, order.astype(np.int32))      # stable ascending
    assert np.array_equal(r.sorted_keys, codes[order])
    c = orc.canonical(r.nodes, r.woop, r.tri_index)
    assert np.array_equal(c.inner, _py_range_tree(r.sorted_keys, leaf))
    assert np.array_equal(c.tris, r.sorted_idx)                        # leaves hold idx[s..e) in sorted order
    assert c.leaf_sizes.sum() == n
    assert r.num_nodes == len(c.inner) and r.num_leaves == len(c.leaf_sizes)


def test_boxes_contain_triangles_with_epsilon(orc):
    verts, tris = scenes.room(3_000, seed=8)
    lo, hi = scenes.bbox(verts)
    eps = np.float32(0.001)
    r = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=8, epsilon=float(eps))
    c = orc.canonical(r.nodes, r.woop, r.tri_index)
    # root child boxes together cover the scene inflated by eps
    b = c.boxes[0]
    rlo = np.array([min(b[0], b[4]), min(b[2], b[6]), min(b[8], b[10])])
    rhi = np.array([max(b[1], b[5]), max(b[3], b[7]), max(b[9], b[11])])
    used = verts[np.unique(tris)]
    assert np.allclose(rlo, used.min(0) - eps, atol=1e-6) and np.allclose(rhi, used.max(0) + eps, atol=1e-6)
    # tracing the LBVH gives the same hits as brute force
    from ntrace_b200 import camera
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 96, 72), 96, 72, cam.far)
    got = orc.compact_trace(r.nodes, r.woop, r.tri_index, rays, True)
    ref = orc.brute_trace(verts, tris, rays, True)
    assert (got[:, 0] == ref[:, 0]).mean() >= 0.999


@pytest.mark.parametrize("bits", [4, 2])
def test_hlbvh_is_a_valid_tree_with_better_or_similar_sah(orc, bits):
    verts, tris = scenes.room(20_000, seed=7)
    lo, hi = scenes.bbox(verts)
    l = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=8)
    h = orc.lbvh_build(verts, tris, lo, hi, hlbvh=True, hlbvh_bits=bits, leaf_size=8)
    assert h.num_clusters > 1
    ch = orc.canonical(h.nodes, h.woop, h.tri_index)
    assert sorted(ch.tris.tolist()) == list(range(len(tris)))        # every triangle exactly once
    assert ch.leaf_sizes.max() <= 8
    sl, sh = orc.compact_sah(l.nodes, l.woop), orc.compact_sah(h.nodes, h.woop)
    assert sh["sah"] <= sl["sah"] * 1.10
    from ntrace_b200 import camera
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 64, 48), 64, 48, cam.far)
    a = orc.compact_trace(l.nodes, l.woop, l.tri_index, rays, True)
    b = orc.compact_trace(h.nodes, h.woop, h.tri_index, rays, True)
    assert (a[:, 0] == b[:, 0]).mean() >= 0.999
