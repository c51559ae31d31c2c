"""Benchmark driver with the reference's loop, knobs and log records.

Mirror of ``FW::runBenchmark`` (src/rt/App.cpp:842-1007) and the option plumbing of ``FW::init``
(src/rt/App.cpp:1011-1083): for every kernel x ray type x camera: ``setParams`` -> ``beginFrame`` ->
``getTotalNumRays`` -> ``while nextBatch(): traceBatch (+ warmupRepeats, measureRepeats timed)``; per (kernel,
ray type) it writes the ``#SUM_RENDER_TIME`` / ``#SUM_RENDER_KRAYS`` records that ``tests/table-tests.sh`` of the
reference greps, then prints the summary table.  (The reference's "Mrays" variable holds Krays/s, App.cpp:969-975;
the record keeps that unit, the printed table adds true Mrays/s.)

    python -m ntrace_b200.app config.conf -DRenderer.builder=HLBVH -DRenderer.rayType=primary;AO;diffuse \\
           -DBenchmark.scene=synthetic:conference -DBenchmark.camera=conference

``Benchmark.scene``: a Wavefront ``.obj`` / ``.ntmesh`` path, or ``synthetic:<sibenik|conference|fairyforest|sanmiguel>``.
``Benchmark.camera``: reference camera signatures separated by ';', or the names of the built-in ones.
"""
from __future__ import annotations

import sys

from . import camera as _camera
from . import host, mesh_io, scenes
from .environment import Environment

RAY_TYPE_NAMES = {"primary": host.RayType_Primary, "ao": host.RayType_AO, "diffuse": host.RayType_Diffuse}
DEFAULT_KERNEL = "b200_persistent_speculative_while_while"


def load_scene(spec: str):
    if spec.startswith("synthetic:"):
        parts = spec.split(":")
        if parts[1] == "room" and len(parts) == 4:                     # synthetic:room:<numTris>:<seed>
            v, t = scenes.room(int(parts[2]), int(parts[3]))
            return v, t, "conference"
        v, t, cam = scenes.config_scene(parts[1])
        return v, t, cam
    if spec.endswith(".ntmesh"):
        v, t = mesh_io.load_ntmesh(spec)
    else:
        v, t = mesh_io.load_obj(spec)
    return v, t, None


def parse_cameras(spec: str, fallback):
    cams = []
    for c in [x for x in (spec or "").split(";") if x.strip()]:
        c = c.strip()
        cams.append(_camera.named_camera(c) if c in _camera.SIGNATURES else _camera.decode_signature(c))
    if not cams and fallback:
        cams.append(_camera.named_camera(fallback))
    if not cams:
        raise host.NtError("Benchmark.camera is empty")
    return cams


def run_benchmark(env: Environment, out=sys.stdout, device: int = 0):
    host.init(device)
    w, h = env.GetInt("App.frameWidth"), env.GetInt("App.frameHeight")
    if env.Has("Renderer.dataStructure") and env.GetString("Renderer.dataStructure") != "BVH":
        raise host.NtError("Incorrect data structure type!  (only Renderer.dataStructure=BVH is on this path)")
    builder = env.GetString("Renderer.builder") if env.Has("Renderer.builder") else "HLBVH"
    verts, tris, scene_cam = load_scene(env.GetString("Benchmark.scene"))
    cameras = parse_cameras(env.GetString("Benchmark.camera") if env.Has("Benchmark.camera") else "", scene_cam)
    kernels = [k for k in (env.GetString("Benchmark.kernel") if env.Has("Benchmark.kernel") else DEFAULT_KERNEL).split(";") if k]
    ray_types = [r.strip() for r in (env.GetString("Renderer.rayType") if env.Has("Renderer.rayType") else "primary").replace(" ", ";").split(";") if r.strip()]
    for r in ray_types:
        if r.lower() not in RAY_TYPE_NAMES:
            raise host.NtError(f"Unsupported ray type {r}")
    warm, meas = env.GetInt("Benchmark.warmupRepeats"), env.GetInt("Benchmark.measureRepeats")
    # NEW knob (default off = the reference's loop): whole frames are repeated instead of single batches, the batches of a frame are
    # queued back to back on alternating buffers (Renderer.setPipelined) and the device time of the whole batch loop is what is summed
    pipelined = env.Has("Benchmark.pipelined") and env.GetBool("Benchmark.pipelined")
    frame_launch = env.Has("Benchmark.frameLaunch") and env.GetBool("Benchmark.frameLaunch")

    print(f'Running benchmark for "{env.GetString("Benchmark.scene")}".\n', file=out)
    scene = host.Scene(verts, tris)
    leaf = env.GetInt("HLBVH.leafSize")
    host.capi.bvh_set_collapse(1 if env.GetBool("HLBVH.collapse") else 0, leaf)
    host.capi.raygen_set_order(1 if env.GetBool("Raygen.coherentOrder") else 0)          # NEW knob (default: the reference's slot order)
    # Renderer.cacheDataStructure + Benchmark.cachePath (new; the reference hard-codes "bvhcache"): files named <hash>_<builder>.dat
    cache_path = env.GetString("Benchmark.cachePath") if env.Has("Benchmark.cachePath") and env.GetBool("Renderer.cacheDataStructure") else None
    renderer = host.Renderer(host.BuildSettings(builder=builder, hlbvh=host.HLBVHParams(True, env.GetInt("HLBVH.bits"), leaf, 0.001), cachePath=cache_path or None,
                                                dataStructure=env.GetString("Renderer.dataStructure") if env.Has("Renderer.dataStructure") else "BVH"))
    renderer.setScene(scene)
    stats = open(env.GetString("App.stats"), "a")
    results = []
    for kernel in kernels:
        for rt in ray_types:
            total_rays, total_time = 0, 0.0
            for ci, cam in enumerate(cameras):
                print(f"{kernel}, {rt}, camera {ci}...", file=out)
                renderer.setParams(host.RendererParams(kernelName=kernel, rayType=RAY_TYPE_NAMES[rt.lower()], numSamples=env.GetInt("Renderer.samples"),
                                                       aoRadius=env.GetFloat("Raygen.aoRadius"), sortSecondary=env.GetBool("Renderer.sortRays")))
                renderer.setPipelined(pipelined)
                if frame_launch:
                    # NEW knob Benchmark.frameLaunch: every batch of the frame is generated into its own buffer, then ONE persistent launch
                    # traces them (Renderer.prepareFrame / traceFrame); the kernel seconds of that launch are what is summed
                    renderer.setPipelined(False)
                    renderer.beginFrame(cam, w, h)
                    total_rays += renderer.getTotalNumRays() * meas
                    renderer.prepareFrame()
                    for _ in range(1 + warm):
                        renderer.traceFrame()
                    for _ in range(meas):
                        total_time += renderer.traceFrame()
                    continue
                if pipelined:
                    for rep in range(warm + meas):
                        renderer.beginFrame(cam, w, h)
                        if rep == 0:
                            total_rays += renderer.getTotalNumRays() * meas
                        renderer.beginTiming()
                        while renderer.nextBatch():
                            renderer.traceBatch()
                        sec = renderer.endTiming()
                        if rep >= warm:
                            total_time += sec
                    continue
                renderer.beginFrame(cam, w, h)
                total_rays += renderer.getTotalNumRays() * meas
                while renderer.nextBatch():
                    renderer.traceBatch()
                    for _ in range(warm):
                        renderer.traceBatch()
                    for _ in range(meas):
                        total_time += renderer.traceBatch()
            krays = total_rays / total_time * 1.0e-3 if total_time > 0 else 0.0
            results.append(krays)
            stats.write(f"#SUM_RENDER_TIME\n{total_time:g}\n#SUM_RENDER_KRAYS\n{krays:g}\n")        # pushStat (Defs.hpp:166-172)
    stats.close()
    print("Done.\n", file=out)
    print("%-42s" % "Kernel" + "".join("%-14s" % r for r in ray_types) + "  [Mrays/s]", file=out)
    print("%-42s" % "---" + "".join("%-14s" % "---" for _ in ray_types), file=out)
    for i, k in enumerate(kernels):
        print("%-42s" % k + "".join("%-14.2f" % (results[i * len(ray_types) + j] * 1e-3) for j in range(len(ray_types))), file=out)
    print("%-42s" % "---" + "".join("%-14s" % "---" for _ in ray_types) + "\n", file=out)
    return results


def main(argv=None):
    env = Environment()
    env.Parse(sys.argv[1:] if argv is None else argv, default_env_file=None)
    Environment.SetSingleton(env)
    if not env.GetBool("App.benchmark"):
        raise host.NtError("only App.benchmark=true is supported (the interactive GUI is out of scope)")
    run_benchmark(env)


if __name__ == "__main__":
    main()
