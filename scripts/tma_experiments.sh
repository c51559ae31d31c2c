# dev tool: the two cp.async.bulk (TMA bulk copy) experiments, against the default paths.  Results: profiles/r2_summary.md
echo "== radix tile load: classic four-pass sort, plain loads vs bulk copy"
NT_SORT_ONESWEEP=0 python scripts/sort_variant_bench.py
NT_SORT_ONESWEEP=0 NT_SORT_BULK=1 python scripts/sort_variant_bench.py
echo "== parity of the builder with the bulk tile load"
NT_SORT_ONESWEEP=0 NT_SORT_BULK=1 python -m pytest tests/test_gpu_build.py tests/test_gpu_raysort.py -x -q 2>&1 | tail -2
echo "== warp ray fetch: direct 256-bit loads vs bulk copy into a per-warp staging area"
python scripts/kernel_compare.py --kernels b200_persistent_speculative_while_while --batches 8 --repeats 3 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'
NT_TRACE_BULKRAYS=1 python scripts/kernel_compare.py --kernels b200_persistent_speculative_while_while --batches 8 --repeats 3 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'
echo "== parity of the trace kernel with the bulk ray fetch"
NT_TRACE_BULKRAYS=1 python -m pytest tests/test_gpu_trace.py tests/test_gpu_bench_frame.py -x -q 2>&1 | tail -2
