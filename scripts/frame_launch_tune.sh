# dev tool: trace-kernel knobs re-checked in the steady state of one launch over the frame's batches (scripts/giant_launch_check.py prints
# per-batch synchronous / two-stream / one-launch Mrays/s for AO and diffuse)
run() { echo "== $*"; env "$@" python scripts/giant_launch_check.py 2>&1 | grep -E "^(AO|diffuse)" | sed -E 's/, (two|results).*one launch of [0-9]+ rays/ -> one launch/; s/; results.*//'; }
run NT_X=0
run NT_TRACE_FETCH=12 NT_WIDE_FETCH=12
run NT_TRACE_FETCH=28 NT_WIDE_FETCH=28
run NT_WIDE_NODE_EXIT=4
run NT_WIDE_NODE_EXIT=12
run NT_TRACE_CARVEOUT=10 NT_WIDE_CARVEOUT=10
run NT_TRACE_CARVEOUT=30 NT_WIDE_CARVEOUT=30
run NT_TRACE_SMEM=16 NT_TRACE_CARVEOUT=35
