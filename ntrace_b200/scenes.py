"""Seeded synthetic stand-ins for the benchmark scenes that are absent offline.

The reference benchmarks Conference / Sibenik / Fairy Forest / San Miguel OBJ files that are
not in the checkout (``.MISSING_LARGE_BLOBS``); BASELINE.json asks for "synthetic triangle
soups of the named triangle count and spatial distribution".  Every generator is a pure
function of ``(num_tris, seed)`` using numpy's PCG64 stream, returns an indexed mesh
``(verts float32 [V,3], tris int32 [T,3])`` with exactly ``num_tris`` triangles, and is
evaluated identically by the CPU oracle and the GPU path (both read the same arrays).

The global triangle index is the row of ``tris`` (reference: ``Scene.cpp:101-117``).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------
# small mesh helpers
# ----------------------------------------------------------------------------------------------
def _grid_quad(p0, du, dv, nu, nv):
    """Tessellated parallelogram p0 + u*du + v*dv, nu x nv quads -> (verts, tris), 2*nu*nv triangles."""
    u = np.linspace(0.0, 1.0, nu + 1, dtype=np.float64)
    v = np.linspace(0.0, 1.0, nv + 1, dtype=np.float64)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    verts = (np.asarray(p0, np.float64)[None, None, :] + uu[..., None] * np.asarray(du, np.float64) + vv[..., None] * np.asarray(dv, np.float64))
    verts = verts.reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).reshape(-1)
    b = a + (nv + 1)
    c = b + 1
    d = a + 1
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    return verts.astype(np.float32), tris.astype(np.int32)


def _box(size, sub):
    """Axis-aligned box centred at the origin, every face tessellated sub x sub -> 12*sub^2 triangles."""
    sx, sy, sz = [0.5 * s for s in size]
    faces = [
        ((-sx, -sy, -sz), (0, 2 * sy, 0), (2 * sx, 0, 0)),   # bottom
        ((-sx, -sy, sz), (2 * sx, 0, 0), (0, 2 * sy, 0)),    # top
        ((-sx, -sy, -sz), (2 * sx, 0, 0), (0, 0, 2 * sz)),   # front
        ((-sx, sy, -sz), (0, 0, 2 * sz), (2 * sx, 0, 0)),    # back
        ((-sx, -sy, -sz), (0, 0, 2 * sz), (0, 2 * sy, 0)),   # left
        ((sx, -sy, -sz), (0, 2 * sy, 0), (0, 0, 2 * sz)),    # right
    ]
    vs, ts, ofs = [], [], 0
    for p0, du, dv in faces:
        v, t = _grid_quad(p0, du, dv, sub, sub)
        vs.append(v); ts.append(t + ofs); ofs += len(v)
    return np.concatenate(vs), np.concatenate(ts)


def _merge(parts):
    vs, ts, ofs = [], [], 0
    for v, t in parts:
        if len(t) == 0:
            continue
        vs.append(np.asarray(v, np.float32)); ts.append(np.asarray(t, np.int32) + ofs); ofs += len(v)
    return np.concatenate(vs).astype(np.float32), np.concatenate(ts).astype(np.int32)


def _translate(mesh, d):
    v, t = mesh
    return v + np.asarray(d, np.float32)[None, :], t


def _instances(template, pos, yaw, scale):
    """Instance a template mesh K times: rotate about z by yaw[k], scale by scale[k] (scalar or xyz), move to pos[k]."""
    tv, tt = template
    k = len(pos)
    c, s = np.cos(yaw), np.sin(yaw)
    sc = np.broadcast_to(np.asarray(scale, np.float64).reshape(k, -1), (k, 3)) if np.ndim(scale) else np.full((k, 3), float(scale))
    p = tv[None, :, :].astype(np.float64) * sc[:, None, :]
    x = p[..., 0] * c[:, None] - p[..., 1] * s[:, None]
    y = p[..., 0] * s[:, None] + p[..., 1] * c[:, None]
    verts = np.stack([x, y, p[..., 2]], -1) + np.asarray(pos, np.float64)[:, None, :]
    tris = tt[None, :, :] + (np.arange(k, dtype=np.int64) * len(tv))[:, None, None]
    return verts.reshape(-1, 3).astype(np.float32), tris.reshape(-1, 3).astype(np.int32)


def _chair(sub):
    """A chair-like prop made of 6 tessellated boxes (seat, back, four legs): 72*sub^2 triangles, unit footprint."""
    parts = [
        _translate(_box((0.5, 0.5, 0.06), sub), (0, 0, 0.45)),
        _translate(_box((0.5, 0.06, 0.55), sub), (0, 0.22, 0.75)),
    ]
    for dx in (-0.2, 0.2):
        for dy in (-0.2, 0.2):
            parts.append(_translate(_box((0.05, 0.05, 0.44), sub), (dx, dy, 0.21)))
    return _merge(parts)


def _cylinder(radius, height, nseg, nstack):
    ang = np.linspace(0.0, 2.0 * np.pi, nseg, endpoint=False)
    z = np.linspace(0.0, height, nstack + 1)
    aa, zz = np.meshgrid(ang, z, indexing="ij")
    verts = np.stack([radius * np.cos(aa), radius * np.sin(aa), zz], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nseg), np.arange(nstack), indexing="ij")
    a = (i * (nstack + 1) + j).reshape(-1)
    b = (((i + 1) % nseg) * (nstack + 1) + j).reshape(-1)
    tris = np.concatenate([np.stack([a, b, b + 1], 1), np.stack([a, b + 1, a + 1], 1)], 0)
    return verts.astype(np.float32), tris.astype(np.int32)


def _uv_sphere(nseg, nring):
    th = np.linspace(0.0, np.pi, nring + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, nseg, endpoint=False)
    tt, pp = np.meshgrid(th, ph, indexing="ij")
    verts = np.stack([np.sin(tt) * np.cos(pp), np.sin(tt) * np.sin(pp), np.cos(tt)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nring), np.arange(nseg), indexing="ij")
    a = (i * nseg + j).reshape(-1)
    b = (i * nseg + (j + 1) % nseg).reshape(-1)
    c = a + nseg
    d = b + nseg
    tris = np.concatenate([np.stack([a, c, d], 1), np.stack([a, d, b], 1)], 0)
    return verts.astype(np.float32), tris.astype(np.int32)


def _random_tris(rng, n, lo, hi, edge_lo, edge_hi):
    """n independent small triangles with centroids uniform in [lo,hi] and edge lengths in [edge_lo,edge_hi]."""
    if n <= 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    c = rng.uniform(lo, hi, size=(n, 3))
    e = rng.uniform(edge_lo, edge_hi, size=(n, 1, 1))
    d = rng.normal(size=(n, 3, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True) + 1e-12
    verts = (c[:, None, :] + 0.5 * e * d).reshape(-1, 3).astype(np.float32)
    tris = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    return verts, tris


def _sub_for(budget, per_unit_at_sub1):
    """Largest tessellation level s with per_unit_at_sub1 * s^2 <= budget (at least 1)."""
    return max(1, int(np.floor(np.sqrt(max(budget, 1) / per_unit_at_sub1))))


# ----------------------------------------------------------------------------------------------
# scenes
# ----------------------------------------------------------------------------------------------
def room(num_tris: int, seed: int, wall_frac: float = 0.3, origin=(0.0, 0.0, 0.0), size=(40.0, 20.0, 15.0)):
    """Architectural interior: closed box with tessellated walls, 12 columns and a floor grid of chair-like
    props (clusters of a few hundred small triangles), topped up with small random clutter so the triangle
    count is exact.  ``wall_frac`` is the fraction of triangles spent on the six walls."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ox, oy, oz = origin
    sx, sy, sz = size
    parts = []

    # walls: six faces, tessellation proportional to edge lengths
    wall_budget = int(num_tris * wall_frac)
    area = 2 * (sx * sy + sy * sz + sx * sz)
    dens = np.sqrt(max(wall_budget, 12) / 2.0 / area)      # quads per unit length
    def q(l):
        return max(1, int(l * dens))
    faces = [
        ((ox, oy, oz), (sx, 0, 0), (0, sy, 0), q(sx), q(sy)),
        ((ox, oy, oz + sz), (0, sy, 0), (sx, 0, 0), q(sy), q(sx)),
        ((ox, oy, oz), (0, 0, sz), (sx, 0, 0), q(sz), q(sx)),
        ((ox, oy + sy, oz), (sx, 0, 0), (0, 0, sz), q(sx), q(sz)),
        ((ox, oy, oz), (0, sy, 0), (0, 0, sz), q(sy), q(sz)),
        ((ox + sx, oy, oz), (0, 0, sz), (0, sy, 0), q(sz), q(sy)),
    ]
    for p0, du, dv, nu, nv in faces:
        parts.append(_grid_quad(p0, du, dv, nu, nv))
    used = sum(len(t) for _, t in parts)

    # 12 columns, two rows along the long axis
    col_budget = int(num_tris * 0.08)
    nseg = 24
    nstack = max(1, col_budget // (12 * 2 * nseg))
    col = _cylinder(0.6, sz, nseg, nstack)
    cpos = [(ox + sx * (i + 0.5) / 6.0, oy + sy * fy, oz) for i in range(6) for fy in (0.2, 0.8)]
    parts.append(_instances(col, np.asarray(cpos), np.zeros(12), 1.0))
    used += 12 * len(col[1])

    # chair-like props on a jittered floor grid
    remaining = num_tris - used
    per_chair_target = 300
    sub = _sub_for(per_chair_target, 72)
    chair = _chair(sub)
    per_chair = len(chair[1])
    n_chairs = max(0, int(remaining * 0.95) // per_chair)
    if n_chairs > 0:
        gx = max(1, int(np.ceil(np.sqrt(n_chairs * sx / sy))))
        gy = max(1, int(np.ceil(n_chairs / gx)))
        k = np.arange(n_chairs)
        px = ox + (k % gx + 0.5) / gx * (sx - 2.0) + 1.0 + rng.uniform(-0.15, 0.15, n_chairs)
        py = oy + (k // gx + 0.5) / gy * (sy - 2.0) + 1.0 + rng.uniform(-0.15, 0.15, n_chairs)
        # dense scenes stack props in layers so footprints do not overlap too heavily
        layer = (k // (gx * gy)).astype(np.float64)
        pz = oz + layer * 1.2
        yaw = rng.uniform(0.0, 2.0 * np.pi, n_chairs)
        cell = min((sx - 2.0) / gx, (sy - 2.0) / gy)
        scale = np.clip(cell * 0.8, 0.05, 1.4) * rng.uniform(0.85, 1.15, n_chairs)
        parts.append(_instances(chair, np.stack([px, py, pz], 1), yaw, scale))
        used += n_chairs * per_chair

    # exact top-up: small clutter triangles just above the floor
    rest = num_tris - used
    if rest < 0:
        raise ValueError("room(): triangle budget too small for the fixed structure")
    parts.append(_random_tris(rng, rest, (ox + 0.5, oy + 0.5, oz + 0.02), (ox + sx - 0.5, oy + sy - 0.5, oz + 0.6), 0.05, 0.5))
    verts, tris = _merge(parts)
    assert len(tris) == num_tris
    return verts, tris


def teapot_in_stadium(num_tris: int, seed: int):
    """Highly non-uniform triangle sizes: 70 % of the triangles in 40 Gaussian-placed detailed blobs
    (sigma = 1 % of the extent) on a coarse ground / backdrop that holds the remaining 30 %."""
    rng = np.random.Generator(np.random.PCG64(seed))
    extent = 4.0
    parts = []
    n_detail = int(num_tris * 0.7)
    n_blobs = 40
    per_blob = n_detail // n_blobs
    nseg = max(8, int(np.sqrt(per_blob / 2.0)))
    nring = max(4, per_blob // (2 * nseg))
    blob = _uv_sphere(nseg, nring)
    centres = rng.normal(0.0, 0.01 * extent * 6.0, size=(n_blobs, 3))
    centres[:, 2] = np.abs(centres[:, 2]) * 0.3 + 0.03
    radii = rng.uniform(0.01, 0.04, n_blobs) * extent / 4.0
    parts.append(_instances(blob, centres, rng.uniform(0, 2 * np.pi, n_blobs), radii))
    used = n_blobs * len(blob[1])
    # backdrop "trees": cones/cylinders around the clearing
    n_back = 24
    tree = _cylinder(0.08, 1.2, 12, 4)
    ang = np.linspace(0, 2 * np.pi, n_back, endpoint=False)
    tp = np.stack([1.5 * np.cos(ang), 1.5 * np.sin(ang), np.zeros(n_back)], 1)
    parts.append(_instances(tree, tp, np.zeros(n_back), 1.0))
    used += n_back * len(tree[1])
    # ground: coarse grid using what is left (two triangles per quad), exact top-up with random large tris
    rest = num_tris - used
    g = max(1, int(np.sqrt(rest / 2.0)))
    parts.append(_grid_quad((-extent / 2, -extent / 2, 0.0), (extent, 0, 0), (0, extent, 0), g, g))
    used += 2 * g * g
    parts.append(_random_tris(rng, num_tris - used, (-1.0, -1.0, 0.01), (1.0, 1.0, 0.05), 0.05, 0.3))
    verts, tris = _merge(parts)
    assert len(tris) == num_tris
    return verts, tris


def soup_uniform(num_tris: int, seed: int, clustered: bool = False):
    """Uniform triangle soup in the unit cube, edge length in [0.5, 1.5] * N^(-1/3).  ``clustered`` puts
    half of the triangles into 1 % of the volume (long Morton duplicate runs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = num_tris
    e = float(n) ** (-1.0 / 3.0)
    c = rng.random((n, 3), dtype=np.float32)
    if clustered:
        half = n // 2
        c[:half] = 0.4 + c[:half] * np.float32(0.2154)      # 0.2154^3 ~ 1 % of the volume
    d = rng.standard_normal((n, 3, 3), dtype=np.float32)
    d /= np.linalg.norm(d, axis=2, keepdims=True) + np.float32(1e-12)
    edge = (rng.random((n, 1, 1), dtype=np.float32) + np.float32(0.5)) * np.float32(e)
    verts = (c[:, None, :] + np.float32(0.5) * edge * d).reshape(-1, 3).astype(np.float32)
    tris = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    return verts, tris


def bbox(verts):
    """Mesh bounding box over *vertices* (reference: Scene.cpp:38 / Mesh.cpp:597-619)."""
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    return v.min(0).astype(np.float32), v.max(0).astype(np.float32)


# name -> (generator, camera signature name) for the five BASELINE.json configs
def config_scene(name: str):
    """Scenes of BASELINE.json ``configs`` by short name -> (verts, tris, camera_name)."""
    if name == "sibenik":          # config 0: ~75K tris, CPU correctness reference
        v, t = room(75_000, 1, wall_frac=0.6, origin=(-20.0, -10.0, 0.0))
        return v, t, "sibenik"
    if name == "conference":       # config 1: ~283K tris, primary + AO + diffuse on 1 GPU
        v, t = room(283_000, 2, wall_frac=0.3)
        return v, t, "conference"
    if name == "fairyforest":      # config 2: ~174K tris, incoherent diffuse
        v, t = teapot_in_stadium(174_000, 3)
        return v, t, "fairyforest"
    if name == "sanmiguel":        # config 3: ~10.5M tris, multi-GPU diffuse
        v, t = room(10_500_000, 4, wall_frac=0.2)
        return v, t, "conference"
    if name == "soup50m":          # config 4: 50M-triangle soup, per-frame rebuild
        v, t = soup_uniform(50_000_000, 5)
        return v, t, "soup"
    raise KeyError(name)
