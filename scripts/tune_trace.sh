#!/bin/bash
# dev tool: sweep the trace-kernel tuning knobs on the bench workload (device-timed region only)
for tri in 0 1 2; do for fetch in 12 20 26 31; do
  echo -n "tri=$tri fetch=$fetch: "; NT_TRACE_TRI=$tri NT_TRACE_FETCH=$fetch python bench.py --steps 3 --warmup 2 --profile 2>&1 | tail -1
done; done
