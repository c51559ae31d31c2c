# dev tool: re-check of the trace kernels' tuning knobs with the direction-coherent ray order (nt_raygen_set_order(1)), kernel b200_auto
run() { echo "== $*"; env "$@" python scripts/kernel_compare.py --kernels ${K:-b200_auto} --batches 24 --repeats 3 --raygen-order 1 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'; }
run NT_X=0
run NT_WIDE_NODE_EXIT=0
run NT_WIDE_NODE_EXIT=4
run NT_WIDE_NODE_EXIT=12
run NT_WIDE_NODE_EXIT=16
run NT_WIDE_FETCH=12 NT_TRACE_FETCH=12
run NT_WIDE_FETCH=26 NT_TRACE_FETCH=26
run NT_WIDE_SMEM=16 NT_WIDE_CARVEOUT=40
K=b200_persistent_speculative_while_while,b200_wide4 run NT_X=0
