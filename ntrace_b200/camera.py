"""Camera signatures and the nscreen->world matrix, mirroring the reference host math in fp32.

Reference: ``src/framework/3d/CameraControls.cpp:250-284`` (orientation / worldToCamera),
``:362-395`` (decodeSignature), ``:491-541`` (6-bit float / direction codec),
``src/framework/base/Math.cpp:66-92`` (fitToView, perspective), ``Math.hpp:1024-1045`` (inverse by
cofactors) and ``src/rt/cuda/Renderer.cpp:473-477`` (nscreenToWorld = invert(fitToView * worldToClip)).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

F = np.float32

# src/rt/App.cpp:51-58, config.conf:11
SIGNATURES = {
    "conference": "6omr/04j3200bR6Z/0/3ZEAz/x4smy19///c/05frY109Qx7w////m100",
    "fairyforest": "cIxMx/sK/Ty/EFu3z/5m9mWx/YPA5z/8///m007toC10AnAHx///Uy200",
    "sibenik": "ytIa02G35kz1i:ZZ/0//iSay/5W6Ex19///c/05frY109Qx7w////m100",
    "sanmiguel": "Yciwz1oRQmz/Xvsm005CwjHx/b70nx18tVI7005frY108Y/:x/v3/z100",
    "armadillo": "GBSvz1V04qy/Ju69/21iChCz/idyKy10A0Kfx1pzUoy/DuY2/0aNqY10sZpuu/5/5f/0/",
}


@dataclass
class Camera:
    position: np.ndarray
    forward: np.ndarray
    up: np.ndarray
    fov: float = 73.74
    near: float = 0.01
    far: float = 100.0
    speed: float = 1.0
    keep_aligned: bool = False


class _Cursor:
    def __init__(self, s):
        self.s, self.i = s, 0

    def bits(self) -> int:
        ch = self.s[self.i]
        self.i += 1
        if "/" <= ch <= ":":
            return ord(ch) - ord("/")
        if "A" <= ch <= "Z":
            return ord(ch) - ord("A") + 12
        if "a" <= ch <= "z":
            return ord(ch) - ord("a") + 38
        raise ValueError("CameraControls: Invalid signature!")

    def f32(self) -> np.float32:
        v = 0
        for sh in range(0, 32, 6):
            v |= self.bits() << sh
        return np.array([v & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0]

    def direction(self) -> np.ndarray:
        face = self.bits()
        tuv = np.array([1.0 if (face & 4) == 0 else -1.0, 0.0, 0.0], dtype=F)
        if (face & 8) == 0:
            tuv[1] = self.f32()
            tuv[2] = self.f32()
        tuv = _normalize(tuv)
        axis = face & 3
        if axis == 0:
            return tuv
        if axis == 1:
            return np.array([tuv[2], tuv[0], tuv[1]], dtype=F)
        return np.array([tuv[1], tuv[2], tuv[0]], dtype=F)


def _normalize(v):
    v = np.asarray(v, dtype=F)
    l = F(np.sqrt(F(np.dot(v, v))))
    return (v * (F(1.0) / l if l != 0 else F(0.0))).astype(F)


def decode_signature(sig: str) -> Camera:
    """CameraControls::decodeSignature (CameraControls.cpp:362-395)."""
    c = _Cursor(sig.strip().strip('"').rstrip(","))
    px, py, pz = c.f32(), c.f32(), c.f32()
    fwd, up = c.direction(), c.direction()
    speed, fov, near, far = c.f32(), c.f32(), c.f32(), c.f32()
    keep = c.bits() != 0
    return Camera(np.array([px, py, pz], dtype=F), fwd, up, float(fov), float(near), float(far), float(speed), keep)


def named_camera(name: str) -> Camera:
    return decode_signature(SIGNATURES[name])


def orientation(cam: Camera) -> np.ndarray:
    """CameraControls::getOrientation (CameraControls.cpp:250-257) -> 3x3 with columns (right, up, -forward)."""
    c2 = -_normalize(cam.forward)
    c0 = _normalize(np.cross(np.asarray(cam.up, F), c2).astype(F))
    c1 = _normalize(np.cross(c2, c0).astype(F))
    return np.stack([c0, c1, c2], axis=1).astype(F)


def world_to_camera(cam: Camera) -> np.ndarray:
    o = orientation(cam)
    pos = (o.T @ np.asarray(cam.position, F)).astype(F)
    m = np.eye(4, dtype=F)
    for i in range(3):
        m[i, :3] = o[:, i]
        m[i, 3] = -pos[i]
    return m


def perspective(fov: float, near: float, far: float) -> np.ndarray:
    """Mat4f::perspective (Math.cpp:79-92)."""
    f = F(1.0) / F(np.tan(F(fov) * F(np.pi) / F(360.0)))
    d = F(1.0) / (F(near) - F(far))
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = f
    m[1, 1] = f
    m[2, 2] = (F(near) + F(far)) * d
    m[2, 3] = F(2.0) * F(near) * F(far) * d
    m[3, 2] = F(-1.0)
    return m


def fit_to_view(w: int, h: int) -> np.ndarray:
    """gl->xformFitToView(-1, 2) == Mat4f::fitToView(pos=-1, size=2, viewSize) (Math.cpp:66-75)."""
    view = np.array([w, h], dtype=F)
    s = F(min(view / F(2.0)))
    m = np.eye(4, dtype=F)
    m[0, 0] = F(2.0) / view[0] * s
    m[1, 1] = F(2.0) / view[1] * s
    return m


def invert4(m: np.ndarray) -> np.ndarray:
    """MatrixBase::inverted (Math.hpp:1024-1045): cofactor expansion in fp32."""
    m = np.asarray(m, dtype=F)
    r = np.zeros((4, 4), dtype=F)
    d = F(0.0)
    si = F(1.0)

    def det3(v):
        return (v[0, 0] * v[1, 1] * v[2, 2] - v[0, 0] * v[1, 2] * v[2, 1] + v[1, 0] * v[2, 1] * v[0, 2]
                - v[1, 0] * v[2, 2] * v[0, 1] + v[2, 0] * v[0, 1] * v[1, 2] - v[2, 0] * v[0, 2] * v[1, 1])

    for i in range(4):
        sj = si
        for j in range(4):
            rows = [k for k in range(4) if k != j]
            cols = [l for l in range(4) if l != i]
            sub = m[np.ix_(rows, cols)].astype(F)
            dd = F(det3(sub)) * sj
            r[i, j] = dd
            d = F(d + dd * m[j, i])
            sj = -sj
        si = -si
    rd = F(1.0) / d if d != 0 else F(0.0)
    return (r * rd * F(4.0)).astype(F)


def nscreen_to_world(cam: Camera, w: int, h: int) -> np.ndarray:
    """invert(fitToView * perspective * worldToCamera) (Renderer.cpp:473-477)."""
    clip = (perspective(cam.fov, cam.near, cam.far) @ world_to_camera(cam)).astype(F)
    return invert4((fit_to_view(w, h) @ clip).astype(F))


def look_at(position, target, up=(0.0, 0.0, 1.0), fov=73.74, near=0.01, far=100.0) -> Camera:
    position = np.asarray(position, F)
    fwd = _normalize(np.asarray(target, F) - position)
    return Camera(position, fwd, np.asarray(up, F), fov, near, far)
