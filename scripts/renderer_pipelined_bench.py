"""Dev measurement for VERDICT item 4: the reference-shaped host loops (Python app.run_benchmark and the C++ ntrace_bench) on the bench frame
(Conference stand-in, GPU HLBVH(2) + collapse, 1024x768, 32 spp), synchronous (the reference's loop) vs Benchmark.pipelined=true.
Prints Mrays/s per ray type from the #SUM_RENDER_KRAYS records.  Usage: python scripts/renderer_pipelined_bench.py [out.json]"""
import io
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntrace_b200 import app, mesh_io, scenes  # noqa: E402
from ntrace_b200.environment import Environment  # noqa: E402

COMMON = [f"-DRaygen.coherentOrder={os.environ.get('NT_BENCH_COHERENT', 'true')}", "-DApp.frameWidth=1024", "-DApp.frameHeight=768", "-DBenchmark.camera=conference", "-DBenchmark.warmupRepeats=2", "-DBenchmark.measureRepeats=10",
          "-DRenderer.dataStructure=BVH", "-DRenderer.builder=HLBVH", "-DRenderer.rayType=primary;AO;diffuse", "-DRenderer.samples=32", "-DRenderer.sortRays=false",
          "-DHLBVH.bits=2", "-DHLBVH.collapse=true"]


def krays(path):
    lines = open(path).read().split("\n")
    return [float(lines[i + 1]) for i, l in enumerate(lines) if l == "#SUM_RENDER_KRAYS"]


def main():
    out = {}
    tmp = tempfile.mkdtemp()
    kernel = os.environ.get("NT_BENCH_KERNEL", "b200_auto")          # the bench's default kernel
    COMMON.append(f"-DBenchmark.kernel={kernel}")
    out["kernel"] = kernel
    MODES = (("synchronous", []), ("pipelined", ["-DBenchmark.pipelined=true"]), ("frame_launch", ["-DBenchmark.frameLaunch=true"]))
    for mode, extra in MODES:
        stats = os.path.join(tmp, f"py_{mode}.log")
        env = Environment()
        env.Parse(COMMON + [f"-DApp.stats={stats}", "-DBenchmark.scene=synthetic:conference"] + extra, default_env_file=None)
        app.run_benchmark(env, out=io.StringIO())
        k = krays(stats)
        out["python_app_" + mode] = dict(zip(("primary", "AO", "diffuse"), [x * 1e-3 for x in k]))
    exe = os.path.join(ROOT, "ntrace_b200", "host_cpp", "ntrace_bench")
    if os.path.exists(exe):
        verts, tris, _ = scenes.config_scene("conference")
        mesh = os.path.join(tmp, "conference.ntmesh")
        mesh_io.save_ntmesh(mesh, verts, tris)
        for mode, extra in MODES:
            stats = os.path.join(tmp, f"cpp_{mode}.log")
            r = subprocess.run([exe] + COMMON + [f"-DApp.stats={stats}", f"-DBenchmark.scene={mesh}"] + extra, capture_output=True, text=True)
            if r.returncode != 0:
                out["cpp_error"] = (r.stdout + r.stderr)[-500:]
                break
            out["cpp_ntrace_bench_" + mode] = dict(zip(("primary", "AO", "diffuse"), [x * 1e-3 for x in krays(stats)]))
    # frame rate of the three ray types together, as bench.py counts it: rays / (sum of times); 786 432 primary hits -> 24 x 1 Mi rays per secondary type
    for k, v in list(out.items()):
        if isinstance(v, dict):
            n = {"primary": 786432.0, "AO": 786432.0 * 32, "diffuse": 786432.0 * 32}
            out[k]["frame_mrays"] = sum(n.values()) / sum(n[t] / v[t] for t in n)
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
