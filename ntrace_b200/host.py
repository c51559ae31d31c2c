"""Host-side mirror of the reference's operator interface for the tracing path.

Same class and method names, argument meaning and error behaviour as the reference's C++ host
(``namespace FW``), implemented on top of the C ABI (``capi``) with torch tensors as device
memory.  Nothing here computes on the CPU: ray generation, BVH build and traversal all run in the
CUDA library; this module only owns buffers and sequencing.

Reference interfaces mirrored:
  RayBuffer        src/rt/ray/RayBuffer.hpp:38-195, RayBuffer.cpp:38-62
  CudaBVH / CudaAS src/rt/cuda/CudaBVH.hpp:137-152, CudaBVH.cpp:105-125 (stream ctor / serialize)
  HLBVHBuilder     src/rt/bvh/HLBVH/HLBVHBuilder.hpp:25-41 (HLBVHParams, CudaBVH subclass)
  CudaBVHTracer    src/rt/cuda/CudaVirtualTracer.hpp:11-26, CudaBVHTracer.cpp:52-168
  RayGen           src/rt/ray/RayGen.cpp:45-74 (primary), :198-232 (ao), :582-600 (batching)
  Scene            src/rt/Scene.cpp:38,101-117 (triVtxIndex, vtxPos, triNormal, bbox)
  Renderer         src/rt/cuda/Renderer.cpp:138-305 (setParams/getCudaBVH), :405-579 (frame loop), :676-710
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass, field

import numpy as np

from . import camera as _camera
from . import capi
from .capi import NtError

BVHLayout_Compact = capi.LAYOUT_COMPACT
BVHLayout_Compact2 = capi.LAYOUT_COMPACT2

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as t
        _torch = t
    return _torch


_device_index = None


def init(device: int = 0):
    """CudaModule::staticInit: pick the device for this process (one GPU per process)."""
    global _device_index
    capi.init(device)
    t = torch()
    t.cuda.set_device(device)
    _device_index = device


def device():
    if _device_index is None:
        raise NtError("ntrace_b200.host.init() has not been called")
    return torch().device("cuda", _device_index)


_pipelined = False       # inside Renderer's pipelined batch loop no torch work is queued, and a device-wide sync would defeat the overlap


def _sync():
    # the library runs on its own stream and is synchronous; torch producers must be drained before a call
    if not _pipelined:
        torch().cuda.synchronize()


# --------------------------------------------------------------------------------------------------
class RayBuffer:
    """Rays (N x 8 f32), results (N x 4 i32) and the id<->slot maps; storage only grows."""

    def __init__(self, n: int = 0, closest_hit: bool = True):
        self._size = 0
        self._cap = 0
        self._rays = self._results = self._id2slot = self._slot2id = None
        self._need_closest = closest_hit
        self.resize(n)

    def getSize(self) -> int:
        return self._size

    def resize(self, n: int):
        if n < 0:
            raise NtError("RayBuffer: negative size")
        if n > self._cap or self._rays is None:
            t, dev = torch(), device()
            cap = max(n, 1)
            old = (self._rays, self._results, self._id2slot, self._slot2id)
            self._rays = t.zeros((cap, 8), dtype=t.float32, device=dev)
            self._results = t.zeros((cap, 4), dtype=t.int32, device=dev)
            self._id2slot = t.zeros(cap, dtype=t.int32, device=dev)
            self._slot2id = t.zeros(cap, dtype=t.int32, device=dev)
            if old[0] is not None and self._size:
                for new, o in zip((self._rays, self._results, self._id2slot, self._slot2id), old):
                    new[: self._size].copy_(o[: self._size])
            self._cap = cap
        self._size = n

    def reserve(self, n: int):
        """Grow the storage to n slots without changing the size (so later resizes queue no allocation / fill)."""
        size = self._size
        self.resize(max(n, size))
        self._size = size

    def setNeedClosestHit(self, c: bool):
        self._need_closest = bool(c)

    def getNeedClosestHit(self) -> bool:
        return self._need_closest

    # device views (torch tensors over the first getSize() slots)
    def getRayBuffer(self):
        return self._rays[: self._size]

    def getResultBuffer(self):
        return self._results[: self._size]

    def getIDToSlotBuffer(self):
        return self._id2slot[: self._size]

    def getSlotToIDBuffer(self):
        return self._slot2id[: self._size]

    def mortonSort(self):
        """RayBuffer::mortonSort (RayBuffer.cpp:103-163): reorder rays by Morton key, keep the id<->slot maps consistent."""
        if self._size > 1:
            _sync()
            capi.ray_sort(self.getRayBuffer(), self.getIDToSlotBuffer(), self.getSlotToIDBuffer(), self._size)

    def setRays(self, rays_np):
        """Upload host rays (N x 8 float32); ids become the identity."""
        rays_np = np.ascontiguousarray(rays_np, dtype=np.float32).reshape(-1, 8)
        self.resize(len(rays_np))
        t = torch()
        self._rays[: self._size].copy_(t.from_numpy(rays_np))
        ar = t.arange(self._size, dtype=t.int32, device=device())
        self._id2slot[: self._size].copy_(ar)
        self._slot2id[: self._size].copy_(ar)

    def rays_host(self) -> np.ndarray:
        return self.getRayBuffer().cpu().numpy()

    def results_host(self) -> np.ndarray:
        return self.getResultBuffer().cpu().numpy()


# --------------------------------------------------------------------------------------------------
class Scene:
    """Flat scene buffers the path needs: triVtxIndex, vtxPos, triNormal, bbox over vertices."""

    def __init__(self, verts, tris):
        t = torch()
        self.verts_host = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        self.tris_host = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        self.vtxPos = t.from_numpy(self.verts_host).to(device())
        self.triVtxIndex = t.from_numpy(self.tris_host).to(device())
        self.triNormal = t.empty((len(self.tris_host), 3), dtype=t.float32, device=device())
        _sync()
        capi.tri_normals(self.vtxPos, self.triVtxIndex, self.triNormal)
        self.bboxMin = self.verts_host.min(0).astype(np.float32)
        self.bboxMax = self.verts_host.max(0).astype(np.float32)

    # host views of the three buffers Scene::hash covers here (Scene.cpp:171-179)
    def triVtxIndex_host(self):
        return self.tris_host

    def vtxPos_host(self):
        return self.verts_host

    def triNormal_host(self):
        return self.triNormal.cpu().numpy()

    def getNumTriangles(self) -> int:
        return len(self.tris_host)

    def getNumVertices(self) -> int:
        return len(self.verts_host)

    def getBBox(self):
        return self.bboxMin, self.bboxMax


# --------------------------------------------------------------------------------------------------
class CudaBVH:
    """The three CudaBVH buffers (nodes, triWoop, triIndex) plus the layout tag."""

    def __init__(self, nodes=None, woop=None, tri_index=None, layout: int = BVHLayout_Compact):
        self.layout = layout
        self.nodes = None if nodes is None else np.ascontiguousarray(nodes).view(np.int32).reshape(-1)
        self.woop = None if woop is None else np.ascontiguousarray(woop).view(np.int32).reshape(-1)
        self.tri_index = None if tri_index is None else np.ascontiguousarray(tri_index, dtype=np.int32).reshape(-1)
        self.resident = False        # True when the buffers live only in the library (GPU build)
        self.generation = 0          # nt_bvh_generation() after the build / upload that made this handle's BVH the resident one
        self.gpu_seconds = 0.0

    def getLayout(self) -> int:
        return self.layout

    def getNodeBuffer(self):
        self._materialise()
        return self.nodes

    def getTriWoopBuffer(self):
        self._materialise()
        return self.woop

    def getTriIndexBuffer(self):
        self._materialise()
        return self.tri_index

    def _materialise(self):
        if self.nodes is None and self.resident:
            if self.generation and capi.bvh_generation() != self.generation:
                raise NtError("CudaBVH: this handle's BVH is no longer the resident one (a later build / upload replaced it before it was downloaded)")
            self.nodes, self.woop, self.tri_index, self.layout = capi.bvh_download()

    # bvhcache format: S32 layout, then 3 x (S64 size, bytes)  (CudaBVH.cpp:105-125, Buffer.cpp:349-381)
    def serialize(self, out: io.BufferedIOBase):
        self._materialise()
        out.write(struct.pack("<i", self.layout))
        for b in (self.nodes, self.woop, self.tri_index):
            raw = b.tobytes()
            out.write(struct.pack("<q", len(raw)))
            out.write(raw)

    @staticmethod
    def deserialize(inp: io.BufferedIOBase) -> "CudaBVH":
        (layout,) = struct.unpack("<i", inp.read(4))
        bufs = []
        for _ in range(3):
            (n,) = struct.unpack("<q", inp.read(8))
            bufs.append(np.frombuffer(inp.read(n), dtype=np.int32).copy())
        return CudaBVH(bufs[0], bufs[1], bufs[2], layout)


@dataclass
class HLBVHParams:
    hlbvh: bool = True
    hlbvhBits: int = 4
    leafSize: int = 8
    epsilon: float = 0.001


class HLBVHBuilder(CudaBVH):
    """GPU LBVH / HLBVH build straight into the Compact layout; the result stays on the device."""

    def __init__(self, scene: Scene, params: HLBVHParams = HLBVHParams()):
        super().__init__(layout=BVHLayout_Compact)
        lbvh = (not params.hlbvh) or params.hlbvhBits == 10          # HLBVHBuilder.cpp:44-47
        _sync()
        self.gpu_seconds = capi.bvh_build(capi.BUILDER_LBVH if lbvh else capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex,
                                          scene.bboxMin, scene.bboxMax, params.hlbvhBits, params.leafSize, params.epsilon)
        self.resident = True
        self.generation = capi.bvh_generation()
        self.layout = capi.bvh_sizes()[1]                 # Compact, or Compact2 after nt_bvh_set_build_layout(5)
        self.num_tris = scene.getNumTriangles()

    def getGPUTime(self) -> float:
        return self.gpu_seconds


# --------------------------------------------------------------------------------------------------
class CudaBVHTracer:
    """CudaVirtualTracer for BVHs: setKernel / getDesiredBVHLayout / setBVH / traceBatch."""

    def __init__(self):
        self._bvh = None
        self._kernel = None
        self.setKernel("b200_persistent_speculative_while_while")

    def setKernel(self, name: str):
        if name == self._kernel:
            return
        capi.set_kernel(name)
        self._kernel = name
        self._config = capi.kernel_config()

    def getDesiredBVHLayout(self) -> int:
        return capi.desired_layout()

    def getKernelConfig(self) -> dict:
        return dict(self._config)

    def setBVH(self, bvh: CudaBVH):
        """The library holds ONE resident BVH.  A handle is uploaded unless it already is that BVH (same generation); a device-built
        handle whose BVH was replaced since is re-uploaded from its host copy if it has one, else refused."""
        self._bvh = bvh
        if bvh is None:
            return
        if bvh.generation and capi.bvh_generation() == bvh.generation:
            return
        if bvh.nodes is None:
            if bvh.resident and not bvh.generation:          # adopt-the-resident handle (replicas after a broadcast, bench.py)
                bvh.generation = capi.bvh_generation()
                return
            raise NtError("CudaBVHTracer: this BVH was built on the device and has been replaced by a later build / upload; rebuild it or keep a host copy")
        capi.bvh_upload(bvh.layout, bvh.nodes, bvh.woop, bvh.tri_index)
        bvh.generation = capi.bvh_generation()

    def traceBatch(self, rays: RayBuffer) -> float:
        """Returns the kernel time in seconds (CUDA events around the launch only)."""
        n = rays.getSize()
        if n == 0:
            return 0.0
        if self._bvh is None:
            raise NtError("CudaBVHTracer: No BVH!")
        if self._bvh.getLayout() != self.getDesiredBVHLayout():
            raise NtError("CudaBVHTracer: Incorrect BVH layout!")
        if self._bvh.generation and capi.bvh_generation() != self._bvh.generation:
            self.setBVH(self._bvh)                           # another handle became resident in between: upload again (or refuse)
        _sync()
        return capi.trace_batch(rays.getRayBuffer(), rays.getResultBuffer(), n, rays.getNeedClosestHit())

    def traceBatches(self, batches) -> float:
        """NEW (nt_trace_batches): several RayBuffers with the same closest / any-hit flag in ONE persistent launch; same results per ray as
        traceBatch on each, without the ramp-up and drain of every launch but one.  Returns the kernel time in seconds."""
        batches = [b for b in batches if b.getSize() > 0]
        if not batches:
            return 0.0
        if self._bvh is None:
            raise NtError("CudaBVHTracer: No BVH!")
        if self._bvh.getLayout() != self.getDesiredBVHLayout():
            raise NtError("CudaBVHTracer: Incorrect BVH layout!")
        if len({b.getNeedClosestHit() for b in batches}) != 1:
            raise NtError("CudaBVHTracer: the batches of one launch share the closest / any-hit flag")
        if self._bvh.generation and capi.bvh_generation() != self._bvh.generation:
            self.setBVH(self._bvh)
        _sync()
        return capi.trace_batches([b.getRayBuffer() for b in batches], [b.getResultBuffer() for b in batches], [b.getSize() for b in batches],
                                  batches[0].getNeedClosestHit())


# --------------------------------------------------------------------------------------------------
class RayGen:
    def __init__(self, maxBatchSize: int = 1 << 20):
        self.m_maxBatchSize = maxBatchSize
        self.m_aoStartIdx = 0
        self.m_shadowStartIdx = 0

    def primary(self, orays: RayBuffer, origin, nscreenToWorld, w: int, h: int, maxDist: float, randomSeed: int = 0):
        orays.resize(w * h)
        orays.setNeedClosestHit(True)
        _sync()
        capi.raygen_primary(orays.getRayBuffer(), orays.getIDToSlotBuffer(), orays.getSlotToIDBuffer(), origin, nscreenToWorld,
                            w, h, maxDist, randomSeed)

    def batching(self, numInputRays: int, numSamples: int, startIdx: int, newBatch: bool):
        """RayGen::batching -> (continues, lo, hi, startIdx, newBatch)."""
        if newBatch:
            newBatch = False
            startIdx = 0
        if startIdx == numInputRays:
            return False, 0, 0, startIdx, newBatch
        lo = startIdx
        hi = min(numInputRays, lo + self.m_maxBatchSize // numSamples)
        return True, lo, hi, hi, newBatch

    def ao(self, orays: RayBuffer, irays: RayBuffer, scene: Scene, numSamples: int, maxDist: float, newBatch: bool, randomSeed: int = 0):
        """RayGen::ao -> (generated, newBatch).  Output needs any-hit only; the caller flips it for diffuse."""
        ok, lo, hi, self.m_aoStartIdx, newBatch = self.batching(irays.getSize(), numSamples, self.m_aoStartIdx, newBatch)
        if not ok:
            return False, newBatch
        orays.resize((hi - lo) * numSamples)
        orays.setNeedClosestHit(False)
        _sync()
        capi.raygen_ao(orays.getRayBuffer(), orays.getIDToSlotBuffer(), orays.getSlotToIDBuffer(), irays.getRayBuffer(),
                       irays.getResultBuffer(), scene.triNormal, lo, hi - lo, numSamples, maxDist, randomSeed)
        return True, newBatch


    def shadow(self, orays: RayBuffer, irays: RayBuffer, numSamples: int, lightPos, lightRadius: float, newBatch: bool, randomSeed: int = 0):
        """RayGen::shadow (RayGen.cpp:114-147) -> (generated, newBatch): any-hit rays from the hit points towards a light."""
        ok, lo, hi, self.m_shadowStartIdx, newBatch = self.batching(irays.getSize(), numSamples, self.m_shadowStartIdx, newBatch)
        if not ok:
            return False, newBatch
        orays.resize((hi - lo) * numSamples)
        orays.setNeedClosestHit(False)
        _sync()
        capi.raygen_shadow(orays.getRayBuffer(), orays.getIDToSlotBuffer(), orays.getSlotToIDBuffer(), irays.getRayBuffer(),
                           irays.getResultBuffer(), lo, hi - lo, numSamples, lightPos, lightRadius, randomSeed)
        return True, newBatch


# --------------------------------------------------------------------------------------------------
RayType_Primary, RayType_AO, RayType_Diffuse = "primary", "AO", "diffuse"

# seed used when Raygen.random is false: the reference hashes Random(0).getU32() (a RANROT value);
# both sides of every parity test use this constant instead.
FIXED_AO_SEED = 0x9E3779B9


@dataclass
class RendererParams:
    kernelName: str = "b200_persistent_speculative_while_while"
    rayType: str = RayType_Primary
    numSamples: int = 32
    aoRadius: float = 5.0
    sortSecondary: bool = False


@dataclass
class BuildSettings:
    builder: str = "HLBVH"            # Renderer.builder: HLBVH | LBVH (GPU); prebuilt CudaBVH via setCudaBVH
    hlbvh: HLBVHParams = field(default_factory=HLBVHParams)
    cachePath: str = None             # Renderer.cacheDataStructure: directory of the bvhcache files (reference: "bvhcache"); None = no cache
    dataStructure: str = "BVH"        # Renderer.dataStructure (part of the cache file name)


# ---- bvhcache file names: Renderer::getCudaBVH (Renderer.cpp:173-178) ------------------------------------------------
_MAGIC = 0x9e3779b9


def _jenkins_mix(a, b, c):
    """FW_JENKINS_MIX (Hash.hpp:172-181), 32-bit wrap-around arithmetic."""
    M = 0xffffffff
    a = (a - b - c) & M; a ^= c >> 13
    b = (b - c - a) & M; b ^= (a << 8) & M
    c = (c - a - b) & M; c ^= b >> 13
    a = (a - b - c) & M; a ^= c >> 12
    b = (b - c - a) & M; b ^= (a << 16) & M
    c = (c - a - b) & M; c ^= b >> 5
    a = (a - b - c) & M; a ^= c >> 3
    b = (b - c - a) & M; b ^= (a << 10) & M
    c = (c - a - b) & M; c ^= b >> 15
    return a, b, c


def hash_bits(a, b=None, c=None, d=None, e=0, f=0):
    """FW::hashBits, both overloads (Hash.hpp:183-184): (a[, b[, c]]) and (a, b, c, d[, e[, f]])."""
    M = 0xffffffff
    if d is None:
        b = _MAGIC if b is None else b
        c = 0 if c is None else c
        a, b, c = _jenkins_mix(a & M, b & M, (c + _MAGIC) & M)
        return c
    a, b, c = _jenkins_mix(a & M, b & M, (c + _MAGIC) & M)
    a, b, c = _jenkins_mix((a + d) & M, (b + e) & M, (c + f) & M)
    return c


def _f2b(x: float) -> int:
    return int(np.float32(x).view(np.uint32))


def cache_file_name(scene: "Scene", builder: str, layout: int, cachePath: str = "bvhcache", dataStructure: str = "BVH",
                    minLeaf: int = 1, maxLeaf: int = 1, splitAlpha: float = 1.0e-5) -> str:
    """"<cachePath>/<hash>_<builder>.dat" exactly as Renderer::getCudaBVH forms it (Renderer.cpp:173-178): hashBits(scene hash,
    Platform("GPU") hash with the renderer's leaf preferences (1, 1) (Renderer.cpp:88-89), BuildParams hash (splitAlpha), layout,
    hash of Renderer.dataStructure).  Scene::hash (Scene.cpp:171-179) covers five buffers; this path carries no material colours,
    so those two enter as empty buffers — for a scene WITH materials the reference's name differs, the file format does not."""
    hb = capi.hash_buffer
    empty = hb(b"")
    scene_hash = hash_bits(hb(scene.triVtxIndex_host()), hb(scene.triNormal_host()), empty, empty, hb(scene.vtxPos_host()))
    platform = hash_bits(hb(b"GPU"), _f2b(1.0), _f2b(1.0), hash_bits(1, 1, minLeaf, maxLeaf))              # Platform.hpp:162
    params = hash_bits(_f2b(splitAlpha))                                                                       # BVH.hpp:141-144
    h = hash_bits(scene_hash, platform, params, layout, hb(dataStructure.encode()))
    return "%s/%08x_%s.dat" % (cachePath, h, builder)


class Renderer:
    """Frame/batch loop of the reference Renderer restricted to the tracing path."""
    PREFETCH = 2

    def __init__(self, build: BuildSettings = BuildSettings()):
        self.m_raygen = RayGen(1 << 20)                               # Renderer.cpp:45
        self.m_cudaTracer = CudaBVHTracer()
        self.m_params = RendererParams()
        self.m_build = build
        self.m_scene = None
        self.m_bvh = None
        self.m_primaryRays = RayBuffer()
        # Three secondary buffers.  The reference has one (Renderer.hpp: m_secondaryRays) and traces batch i before it generates
        # batch i+1; in pipelined mode batches are generated up to PREFETCH ahead into the other buffers while batch i is traced, and
        # consecutive launches overlap at their tails (nt_set_deferred(2); the library orders each call behind the launches / the
        # generator that use ITS buffers only).
        self.m_secondary = [RayBuffer(), RayBuffer(), RayBuffer()]
        self.m_secondaryRays = self.m_secondary[0]
        self.m_genIdx = 0
        self.m_queue = []            # pipelined mode: batches generated ahead of the one handed out (at most PREFETCH)
        self.m_genDone = False
        self.m_pipelined = False
        self.m_timing = False
        self.m_cameraFar = 0.0
        self.m_newBatch = True
        self.m_batchRays = None
        self.m_batchStart = 0

    def setScene(self, scene: Scene):
        self.m_scene = scene
        self.m_bvh = None

    def setParams(self, params: RendererParams):
        self.m_params = params
        self.m_cudaTracer.setKernel(params.kernelName)

    def setPipelined(self, on: bool):
        """NEW (no reference counterpart).  False: the reference's loop, every traceBatch() returns its kernel seconds.  True:
        nextBatch() alternates the two secondary buffers and traceBatch() only queues (returns 0.0); bracket the batch loop with
        beginTiming() / endTiming() to get the device seconds of everything queued in between.  Results are bit-identical."""
        self.m_pipelined = bool(on)

    def beginTiming(self):
        global _pipelined
        if self.m_pipelined:
            for rb in self.m_secondary:                               # no allocation (torch fill kernels) inside the loop
                rb.reserve(self.m_raygen.m_maxBatchSize)
            torch().cuda.synchronize()
            capi.set_deferred(2)
            _pipelined = True
        capi.event_record(6)
        self.m_timing = True

    def endTiming(self) -> float:
        """Device seconds since beginTiming() (waits for everything queued)."""
        global _pipelined
        capi.event_record(7)
        sec = capi.event_elapsed(6, 7)
        if self.m_pipelined:
            capi.synchronize()
            capi.set_deferred(0)
            _pipelined = False
        self.m_timing = False
        return sec

    def setCudaBVH(self, bvh: CudaBVH):
        """Use a prebuilt CudaBVH (e.g. a bvhcache file or a CPU-built SplitBVH flattened by the caller)."""
        self.m_bvh = bvh

    def getCudaBVH(self) -> CudaBVH:
        if self.m_bvh is not None:
            return self.m_bvh
        b = self.m_build.builder
        cache = None
        if self.m_build.cachePath is not None:                       # Renderer.cpp:173-191: a cache file that exists is imported
            import os
            cache = cache_file_name(self.m_scene, b, self.m_cudaTracer.getDesiredBVHLayout(), self.m_build.cachePath, self.m_build.dataStructure)
            if os.path.exists(cache):
                with open(cache, "rb") as f:
                    self.m_bvh = CudaBVH.deserialize(f)
                return self.m_bvh
        if b == "HLBVH":
            self.m_bvh = HLBVHBuilder(self.m_scene, self.m_build.hlbvh)     # Renderer.cpp:201-209
        elif b == "LBVH":
            p = HLBVHParams(False, 10, self.m_build.hlbvh.leafSize, self.m_build.hlbvh.epsilon)
            self.m_bvh = HLBVHBuilder(self.m_scene, p)
        else:
            raise NtError(f"Unsupported BVH builder {b}")
        if cache is not None:                                         # Renderer.cpp:293-299
            import os
            os.makedirs(os.path.dirname(cache) or ".", exist_ok=True)
            with open(cache, "wb") as f:
                self.m_bvh.serialize(f)
        return self.m_bvh

    def beginFrame(self, cam: "_camera.Camera", w: int, h: int):
        self.m_cudaTracer.setBVH(self.getCudaBVH())
        n2w = _camera.nscreen_to_world(cam, w, h)
        self.m_raygen.primary(self.m_primaryRays, cam.position, n2w, w, h, cam.far, 0)
        if self.m_params.rayType != RayType_Primary:                   # Renderer.cpp:481-484
            self.m_cudaTracer.traceBatch(self.m_primaryRays)
        self.m_cameraFar = cam.far
        self.m_newBatch = True
        self.m_batchRays = None
        self.m_batchStart = 0
        self.m_queue = []
        self.m_genDone = False

    def _nextBatchPipelined(self) -> bool:
        """Secondary batches, pipelined: keep up to PREFETCH generated batches queued ahead of the one handed out."""
        closest = self.m_params.rayType == RayType_Diffuse
        dist = self.m_cameraFar if closest else self.m_params.aoRadius
        while len(self.m_queue) < self.PREFETCH and not self.m_genDone:
            rb = self.m_secondary[self.m_genIdx % len(self.m_secondary)]
            ok, self.m_newBatch = self.m_raygen.ao(rb, self.m_primaryRays, self.m_scene, self.m_params.numSamples, dist, self.m_newBatch, FIXED_AO_SEED)
            if not ok:
                self.m_genDone = True
                break
            self.m_genIdx += 1
            rb.setNeedClosestHit(closest)
            if self.m_params.sortSecondary:
                rb.mortonSort()
            self.m_queue.append(rb)
        if not self.m_queue:
            return False
        self.m_batchRays = self.m_secondaryRays = self.m_queue.pop(0)
        return True

    def nextBatch(self) -> bool:
        if self.m_batchRays is not None:
            self.m_batchStart += self.m_batchRays.getSize()
        self.m_batchRays = None
        rt = self.m_params.rayType
        if self.m_pipelined and rt in (RayType_AO, RayType_Diffuse):
            return self._nextBatchPipelined()
        if rt == RayType_Primary:
            if not self.m_newBatch:
                return False
            self.m_newBatch = False
            self.m_batchRays = self.m_primaryRays
        elif rt == RayType_AO:
            ok, self.m_newBatch = self.m_raygen.ao(self.m_secondaryRays, self.m_primaryRays, self.m_scene, self.m_params.numSamples,
                                                  self.m_params.aoRadius, self.m_newBatch, FIXED_AO_SEED)
            if not ok:
                return False
            self.m_batchRays = self.m_secondaryRays
        elif rt == RayType_Diffuse:
            ok, self.m_newBatch = self.m_raygen.ao(self.m_secondaryRays, self.m_primaryRays, self.m_scene, self.m_params.numSamples,
                                                  self.m_cameraFar, self.m_newBatch, FIXED_AO_SEED)
            if not ok:
                return False
            self.m_secondaryRays.setNeedClosestHit(True)
            self.m_batchRays = self.m_secondaryRays
        else:
            raise NtError(f"unsupported ray type {rt}")
        # Renderer.cpp:561-562: the reference's condition is a tautology, so primary batches are sorted too
        if self.m_params.sortSecondary:
            self.m_batchRays.mortonSort()
        return True

    def traceBatch(self) -> float:
        if self.m_batchRays is None:
            raise NtError("Renderer: no batch")
        return self.m_cudaTracer.traceBatch(self.m_batchRays)

    def prepareFrame(self) -> int:
        """NEW (no reference counterpart; the reference's Renderer holds ONE secondary RayBuffer, Renderer.hpp).  After beginFrame():
        generate EVERY batch of the frame for the current ray type, each into its own RayBuffer of a grow-only pool -> number of batches.
        Same batches, same rays as the nextBatch() loop produces one after the other."""
        rt = self.m_params.rayType
        if rt == RayType_Primary:
            self.m_frameBatches = [self.m_primaryRays]
            return 1
        if rt not in (RayType_AO, RayType_Diffuse):
            raise NtError(f"unsupported ray type {rt}")
        closest = rt == RayType_Diffuse
        dist = self.m_cameraFar if closest else self.m_params.aoRadius
        if not hasattr(self, "m_framePool"):
            self.m_framePool = []
        batches, new = [], True
        self.m_raygen.m_aoStartIdx = 0
        while True:
            if len(batches) == len(self.m_framePool):
                self.m_framePool.append(RayBuffer())
            rb = self.m_framePool[len(batches)]
            ok, new = self.m_raygen.ao(rb, self.m_primaryRays, self.m_scene, self.m_params.numSamples, dist, new, FIXED_AO_SEED)
            if not ok:
                break
            rb.setNeedClosestHit(closest)
            if self.m_params.sortSecondary:
                rb.mortonSort()
            batches.append(rb)
        self.m_frameBatches = batches
        return len(batches)

    def traceFrame(self) -> float:
        """NEW: the batches prepareFrame() generated, traced by ONE persistent launch (CudaBVHTracer.traceBatches = nt_trace_batches): the same
        results per ray as traceBatch() on each, without the ramp-up and drain of every launch but one.  Returns the kernel seconds."""
        batches = getattr(self, "m_frameBatches", None)
        if not batches:
            raise NtError("Renderer: no frame prepared")
        if len(batches) == 1:
            return self.m_cudaTracer.traceBatch(batches[0])
        return self.m_cudaTracer.traceBatches(batches)

    def getFrameBatches(self):
        return list(getattr(self, "m_frameBatches", []))

    def getTotalNumRays(self) -> int:
        """Rays counted by the benchmark: w*h for primary, primary hits x samples otherwise (Renderer.cpp:676-710)."""
        if self.m_params.rayType == RayType_Primary:
            return self.m_primaryRays.getSize()
        _sync()
        hits = capi.count_hits(self.m_primaryRays.getResultBuffer(), self.m_primaryRays.getSize())
        return hits * self.m_params.numSamples
