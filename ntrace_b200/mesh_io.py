"""Mesh ingestion with the reference's rules, so that the global triangle index (= hit id) matches
(SURVEY.md App. E):

* ``importWavefrontMesh`` (src/framework/io/MeshWavefrontIO.cpp:412-485): one mesh vertex per distinct
  (position, texcoord, normal) index triple in first-use order; negative (relative) indices; polygons are
  fan-triangulated (v0, v[i-1], v[i]); faces are flushed to the current submesh on every ``usemtl`` and at EOF;
  submeshes are created in first-use order of *known* materials; faces with no / unknown material go to one
  default submesh created on first use.
* ``Scene::Scene`` (src/rt/Scene.cpp:101-117) concatenates the submeshes in index order.

Also a trivial binary ``.ntmesh`` container so CPU and GPU harnesses read the same bytes.
"""
from __future__ import annotations

import os
import struct

import numpy as np


def _mtl_names(path):
    names = []
    try:
        with open(path, "r", errors="replace") as f:
            for line in f:
                p = line.split()
                if len(p) >= 2 and p[0] == "newmtl":
                    names.append(line.strip()[len("newmtl"):].strip())
    except OSError:
        pass
    return names


def load_obj(path: str):
    """-> (verts float32 [V,3], tris int32 [T,3]) in the reference's vertex and triangle order."""
    positions, n_tex, n_nrm = [], 0, 0
    vert_hash, verts = {}, []
    submeshes = []                 # list of lists of triangles
    materials = {}                 # name -> submesh index or -1 (known, unused yet)
    submesh, default_submesh = -1, -1
    index_tmp = []
    dirname = os.path.dirname(path)

    def flush():
        nonlocal index_tmp
        if submesh != -1:
            submeshes[submesh].extend(index_tmp)
        index_tmp = []

    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.strip()
            if not line or line[0] == "#":
                continue
            if line.startswith("v "):
                p = line.split()
                positions.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vt "):
                n_tex += 1
            elif line.startswith("vn "):
                n_nrm += 1
            elif line.startswith("f "):
                tmp = []
                sizes = (len(positions), n_tex, n_nrm)
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    ptn = [0, 0, 0]
                    for i in range(3):
                        v = int(parts[i]) if i < len(parts) and parts[i] not in ("", None) else 0
                        v = v + sizes[i] if v < 0 else v - 1
                        if v < 0 or v >= sizes[i]:
                            v = -1
                        ptn[i] = v
                    key = tuple(ptn)
                    idx = vert_hash.get(key)
                    if idx is None:
                        idx = len(verts)
                        vert_hash[key] = idx
                        verts.append(positions[ptn[0]] if ptn[0] != -1 else (0.0, 0.0, 0.0))
                    tmp.append(idx)
                if submesh == -1:
                    if default_submesh == -1:
                        default_submesh = len(submeshes)
                        submeshes.append([])
                    submesh = default_submesh
                for i in range(2, len(tmp)):
                    index_tmp.append((tmp[0], tmp[i - 1], tmp[i]))
            elif line.startswith("usemtl "):
                name = line[len("usemtl"):].strip()
                if submesh != -1:
                    flush()
                    submesh = -1
                if name in materials:
                    if materials[name] == -1:
                        materials[name] = len(submeshes)
                        submeshes.append([])
                    submesh = materials[name]
                    index_tmp = []
            elif line.startswith("mtllib "):
                for n in _mtl_names(os.path.join(dirname, line[len("mtllib"):].strip())):
                    materials.setdefault(n, -1)
    flush()
    tris = [t for sm in submeshes for t in sm]
    return np.asarray(verts, dtype=np.float32).reshape(-1, 3), np.asarray(tris, dtype=np.int32).reshape(-1, 3)


_MAGIC = b"NTMESH1\0"


def save_ntmesh(path: str, verts, tris):
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<qq", len(verts), len(tris)))
        f.write(verts.tobytes())
        f.write(tris.tobytes())


def load_ntmesh(path: str):
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError("not an .ntmesh file")
        nv, nt = struct.unpack("<qq", f.read(16))
        verts = np.frombuffer(f.read(nv * 12), dtype=np.float32).reshape(nv, 3).copy()
        tris = np.frombuffer(f.read(nt * 12), dtype=np.int32).reshape(nt, 3).copy()
    return verts, tris
