"""Structural invariants of the restated CPU builders (SAHBVHBuilder / SplitBVHBuilder) and createCompact."""
import numpy as np
import pytest

from ntrace_b200 import scenes


@pytest.fixture(scope="module")
def mesh():
    return scenes.room(6_000, seed=3, wall_frac=0.3)


def _walk(nodes, woop, idx):
    """DFS over a Compact tree -> list of (addr, child boxes, child links), leaf tri lists."""
    n = nodes.reshape(-1, 16)
    f = n.view(np.float32)
    seen_tris, inner = [], []
    stack = [0]
    while stack:
        a = stack.pop()
        row = a // 64
        inner.append(row)
        for c in (n[row, 12], n[row, 13]):
            if c >= 0:
                stack.append(int(c))
            else:
                t = ~int(c)
                while woop[t * 4] != np.int32(-2147483648):
                    seen_tris.append(int(idx[t]))
                    t += 3
    return inner, seen_tris, f


@pytest.mark.parametrize("builder", ["sah", "split"])
def test_compact_tree_is_complete_and_boxes_nest(orc, mesh, builder):
    verts, tris = mesh
    bvh = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if builder == "split" else orc.BUILDER_SAH, 1, 1)
    st = bvh.stats()
    nodes, woop, idx = bvh.compact()
    inner, seen, f = _walk(nodes, woop, idx)
    assert len(inner) == st.num_inner == len(nodes) // 16
    assert set(seen) == set(range(len(tris)))                    # every triangle is referenced by >= 1 leaf
    if builder == "sah":
        assert len(seen) == len(tris) and st.duplicates == 0
    else:
        assert len(seen) == len(tris) + st.duplicates
    assert st.num_leaf == st.num_inner + 1
    # child boxes contain the grandchildren boxes
    n = nodes.reshape(-1, 16)
    for row in inner[:2000]:
        for k, c in enumerate((n[row, 12], n[row, 13])):
            if c < 0:
                continue
            lo = np.array([f[row, 4 * k + 0], f[row, 4 * k + 2], f[row, 8 + 2 * k]])
            hi = np.array([f[row, 4 * k + 1], f[row, 4 * k + 3], f[row, 9 + 2 * k]])
            cr = c // 64
            clo = np.minimum([f[cr, 0], f[cr, 2], f[cr, 8]], [f[cr, 4], f[cr, 6], f[cr, 10]])
            chi = np.maximum([f[cr, 1], f[cr, 3], f[cr, 9]], [f[cr, 5], f[cr, 7], f[cr, 11]])
            assert (clo >= lo).all() and (chi <= hi).all()


def test_sah_metric_and_split_quality(orc, mesh):
    verts, tris = mesh
    sah = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1)
    spl = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT, 1, 1)
    s1, s2 = sah.stats(), spl.stats()
    # the flat-tree SAH evaluator reproduces BVHNode::computeSubtreeProbabilities on both trees
    for bvh, st in ((sah, s1), (spl, s2)):
        n, w, _ = bvh.compact()
        cs = orc.compact_sah(n, w)
        assert abs(cs["sah"] - st.sah) <= 1e-4 * st.sah
        assert cs["num_inner"] == st.num_inner and cs["num_tris"] == st.num_tris
    assert s2.sah <= s1.sah * 1.02                                # spatial splits should not make it worse
    # root is always an inner node; leaf preference (1,1) -> one triangle per leaf
    assert s1.num_tris == len(tris) and s1.num_leaf == len(tris)


def test_leaf_preferences_and_degenerates(orc):
    verts, tris = scenes.room(3_000, seed=4)
    # add degenerate triangles (zero-area: two identical vertices, and a point) -> dropped by both builders
    v = np.vstack([verts, [[1, 1, 1], [1, 1, 1], [2, 1, 1], [3, 3, 3]]]).astype(np.float32)
    base = len(verts)
    t = np.vstack([tris, [[base, base + 1, base + 2], [base + 3, base + 3, base + 3]]]).astype(np.int32)
    for b in (orc.BUILDER_SAH, orc.BUILDER_SPLIT):
        bvh = orc.CpuBVH(v, t, b, 1, 8)
        st = bvh.stats()
        nodes, woop, idx = bvh.compact()
        _, seen, _ = _walk(nodes, woop, idx)
        assert len(t) - 1 not in seen                             # the point triangle (all extents zero) is removed
        assert st.num_leaf <= len(t)
    big = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 8).stats()
    one = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1).stats()
    assert big.num_leaf < one.num_leaf
