"""The benchmark driver (mirror of App::runBenchmark) end to end on a small scene: knobs from an environment file +
-D overrides, the reference's stats records, ray accounting of Renderer::getTotalNumRays."""
import io

import pytest

pytestmark = pytest.mark.gpu


def test_run_benchmark_writes_reference_stat_records(gpu_host, tmp_path):
    from ntrace_b200 import app
    from ntrace_b200.environment import Environment
    conf = tmp_path / "config.conf"
    stats = tmp_path / "stats.log"
    conf.write_text(f"""
App {{ benchmark true
stats {stats}
frameWidth 256
frameHeight 192 }}
Benchmark {{ scene synthetic:room:8000:3
camera conference
warmupRepeats 1
measureRepeats 2 }}
Renderer {{ dataStructure BVH
builder HLBVH
rayType primary
samples 4
sortRays false }}
""")
    env = Environment()
    env.Parse([str(conf), "-DRenderer.rayType=primary;AO;diffuse", "-DBenchmark.kernel=b200_persistent_speculative_while_while;fermi_speculative_while_while"])
    out = io.StringIO()
    res = app.run_benchmark(env, out=out)
    assert len(res) == 6 and all(r > 0 for r in res)
    lines = stats.read_text().split("\n")
    assert lines.count("#SUM_RENDER_TIME") == 6 and lines.count("#SUM_RENDER_KRAYS") == 6
    assert "b200_persistent_speculative_while_while" in out.getvalue() and "Done." in out.getvalue()
    # LBVH builder name and the kd-tree data structure of the shipped config are handled / rejected as the reference does
    env.Set("Renderer.builder", "LBVH")
    env.Set("Renderer.rayType", "primary")
    env.Set("Benchmark.kernel", "b200_speculative_while_while")
    assert len(app.run_benchmark(env, out=io.StringIO())) == 1
    env.Set("Renderer.dataStructure", "KDTree")
    with pytest.raises(gpu_host.NtError, match="Incorrect data structure type"):
        app.run_benchmark(env, out=io.StringIO())
    env.Set("Renderer.dataStructure", "BVH")
    env.Set("Renderer.builder", "SplitBVH")
    with pytest.raises(gpu_host.NtError, match="Unsupported BVH builder"):
        app.run_benchmark(env, out=io.StringIO())


def test_pipelined_renderer_loop_is_bit_identical_to_the_synchronous_loop(gpu_host):
    """Renderer.setPipelined(True): two alternating secondary buffers, queued launches that overlap at their tails, ray generation
    ordered only behind the launches that use its buffers.  Every batch's results must equal the reference-style loop's, bit for bit."""
    import numpy as np
    from ntrace_b200 import camera, scenes
    verts, tris = scenes.room(20_000, seed=9, wall_frac=0.3)
    cam = camera.named_camera("conference")
    scene = gpu_host.Scene(verts, tris)
    w, h = 320, 240

    def frame(pipelined, ray_type):
        r = gpu_host.Renderer(gpu_host.BuildSettings(builder="HLBVH"))
        r.m_raygen = gpu_host.RayGen(1 << 16)                      # small batches: many launches in flight
        r.setScene(scene)
        r.setParams(gpu_host.RendererParams(kernelName="b200_persistent_speculative_while_while", rayType=ray_type, numSamples=8, aoRadius=5.0, sortSecondary=False))
        r.setPipelined(pipelined)
        r.beginFrame(cam, w, h)
        total = r.getTotalNumRays()
        out = []
        r.beginTiming()
        while r.nextBatch():
            sec = r.traceBatch()
            assert (sec == 0.0) == pipelined
            if not pipelined:
                out.append(r.m_batchRays.results_host().copy())
            else:
                out.append(r.m_batchRays)                           # results are read after the loop; only the last two buffers survive
        t = r.endTiming()
        assert t > 0.0
        if pipelined:
            out = [b.results_host().copy() for b in out[-3:]]
        return total, out

    for rt in (gpu_host.RayType_AO, gpu_host.RayType_Diffuse):
        n_sync, res_sync = frame(False, rt)
        n_pipe, res_pipe = frame(True, rt)
        assert n_sync == n_pipe and len(res_sync) >= 4
        # the pipelined loop cycles three buffers: its last three batches are still in them
        for k in (1, 2, 3):
            assert np.array_equal(res_pipe[-k][: len(res_sync[-k])], res_sync[-k]), k


def test_run_benchmark_pipelined_knob(gpu_host, tmp_path):
    from ntrace_b200 import app
    from ntrace_b200.environment import Environment
    stats = tmp_path / "stats.log"
    env = Environment()
    env.Parse([f"-DApp.stats={stats}", "-DApp.frameWidth=256", "-DApp.frameHeight=192", "-DBenchmark.scene=synthetic:room:8000:3", "-DBenchmark.camera=conference",
               "-DBenchmark.warmupRepeats=1", "-DBenchmark.measureRepeats=2", "-DRenderer.dataStructure=BVH", "-DRenderer.builder=HLBVH",
               "-DRenderer.rayType=primary;AO;diffuse", "-DRenderer.samples=4", "-DRenderer.sortRays=false", "-DBenchmark.pipelined=true",
               "-DHLBVH.bits=2", "-DHLBVH.collapse=true"], default_env_file=None)
    try:
        res = app.run_benchmark(env, out=io.StringIO())
    finally:
        gpu_host.capi.bvh_set_collapse(0, 0)
    assert len(res) == 3 and all(r > 0 for r in res)
    assert stats.read_text().split("\n").count("#SUM_RENDER_KRAYS") == 3


def test_frame_launch_is_bit_identical_to_the_batch_loop(gpu_host):
    """Renderer.prepareFrame / traceFrame (all batches of the frame in one persistent launch) against the reference-shaped batch loop."""
    import numpy as np
    from ntrace_b200 import camera, scenes
    verts, tris = scenes.room(12_000, seed=5, wall_frac=0.3)
    scene = gpu_host.Scene(verts, tris)
    cam = camera.named_camera("conference")
    w, h = 320, 240
    for rt in (gpu_host.RayType_AO, gpu_host.RayType_Diffuse):
        for kernel in ("b200_persistent_speculative_while_while", "b200_auto"):
            r = gpu_host.Renderer(gpu_host.BuildSettings(builder="HLBVH"))
            r.m_raygen = gpu_host.RayGen(1 << 16)
            r.setScene(scene)
            r.setParams(gpu_host.RendererParams(kernelName=kernel, rayType=rt, numSamples=8, aoRadius=5.0, sortSecondary=False))
            r.beginFrame(cam, w, h)
            want = []
            while r.nextBatch():
                assert r.traceBatch() > 0.0
                want.append(r.m_batchRays.results_host().copy())
            r.beginFrame(cam, w, h)
            assert r.prepareFrame() == len(want) >= 4
            assert r.traceFrame() > 0.0
            got = [b.results_host() for b in r.getFrameBatches()]
            for a, b in zip(want, got):
                assert np.array_equal(a, b)
    gpu_host.capi.set_kernel("b200_persistent_speculative_while_while")


def test_run_benchmark_frame_launch_knob(gpu_host, tmp_path):
    from ntrace_b200 import app
    from ntrace_b200.environment import Environment
    stats = tmp_path / "stats.log"
    env = Environment()
    env.Parse([f"-DApp.stats={stats}", "-DApp.frameWidth=256", "-DApp.frameHeight=192", "-DBenchmark.scene=synthetic:room:8000:3", "-DBenchmark.camera=conference",
               "-DBenchmark.warmupRepeats=1", "-DBenchmark.measureRepeats=2", "-DRenderer.dataStructure=BVH", "-DRenderer.builder=HLBVH",
               "-DRenderer.rayType=primary;AO;diffuse", "-DRenderer.samples=4", "-DRenderer.sortRays=false", "-DBenchmark.frameLaunch=true",
               "-DBenchmark.kernel=b200_auto", "-DRaygen.coherentOrder=true", "-DHLBVH.bits=2", "-DHLBVH.collapse=true"], default_env_file=None)
    try:
        res = app.run_benchmark(env, out=io.StringIO())
    finally:
        gpu_host.capi.bvh_set_collapse(0, 0)
        gpu_host.capi.raygen_set_order(0)
        gpu_host.capi.set_kernel("b200_persistent_speculative_while_while")
    assert len(res) == 3 and all(r > 0 for r in res)
    assert stats.read_text().split("\n").count("#SUM_RENDER_KRAYS") == 3
