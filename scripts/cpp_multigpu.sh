#!/bin/bash
# C++ host (ntrace_bench) on N GPUs of one box, one process per GPU, no Python / MPI in the ranks: the NCCL unique id travels through a file.
#   scripts/cpp_multigpu.sh <N> <scene: conference|sanmiguel> [extra -D options]
# Prints rank 0's output (summary table + broadcast time) and leaves stats in gpurun_out/cpp_multigpu_<scene>_n<N>.log
set -e
N=${1:-2}; SCENE=${2:-conference}; shift 2 || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
python - <<PY
import sys
sys.path.insert(0, "$ROOT")
from ntrace_b200 import mesh_io, scenes
v, t, _ = scenes.config_scene("$SCENE")
mesh_io.save_ntmesh("$TMP/scene.ntmesh", v, t)
PY
mkdir -p "$ROOT/gpurun_out"
OUT="$ROOT/gpurun_out/cpp_multigpu_${SCENE}_n${N}.log"
rm -f "$OUT"
ARGS="-DApp.stats=$OUT -DBenchmark.scene=$TMP/scene.ntmesh -DBenchmark.camera=conference -DBenchmark.warmupRepeats=2 -DBenchmark.measureRepeats=5 -DRenderer.dataStructure=BVH -DRenderer.builder=HLBVH -DRenderer.rayType=diffuse -DRenderer.samples=32 -DRenderer.sortRays=false -DHLBVH.bits=4 -DHLBVH.collapse=true -DBenchmark.pipelined=true -DBenchmark.commFile=$TMP/nccl_id $*"
pids=()
for ((r = 1; r < N; r++)); do
  RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r "$ROOT/ntrace_b200/host_cpp/ntrace_bench" $ARGS > "$TMP/rank$r.out" 2>&1 &
  pids+=($!)
done
RANK=0 WORLD_SIZE=$N LOCAL_RANK=0 "$ROOT/ntrace_b200/host_cpp/ntrace_bench" $ARGS
for p in "${pids[@]}"; do wait $p; done
cat "$OUT"
rm -rf "$TMP"
