#!/bin/bash
# does the clocks sampler (or anything else outside the device loop) cost time inside the timed region?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "X=1" "BENCH_NO_CLOCKS=1" "BENCH_CLOCKS_MS=100" "X=2"; do
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'device loops', [round(x,1) for x in d['loop_ms_device']], d.get('clocks'))"
done
