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
