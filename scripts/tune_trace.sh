#!/bin/bash
# dev tool: sweep the trace-kernel tuning knobs on the bench workload (device-timed region only)
for smem in 0 8 16; do for carve in 0 10 20; do
  echo -n "smem=$smem carve=$carve: "; NT_TRACE_SMEM=$smem NT_TRACE_CARVEOUT=$carve python bench.py --steps 3 --warmup 2 --profile 2>&1 | tail -1
done; done
