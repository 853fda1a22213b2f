#!/bin/bash
# A/B within one box of an env-switched variant: usage run_r02_ab3.sh VAR  (runs VAR=0, VAR=1, twice)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=${1:-DL_V3_STREAM_TASKS}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden and not T500 or config_shapes or odd or sweep or nan or equivariance" 2>&1 | tail -3
for ov in ${VALS:-0 1 0 1}; do
  env $VAR=$ov timeout 300 python bench.py --steps 3 --warmup 3 --T 100 --no-e2e --no-cpu-baseline > gpurun_out/ab_$ov.json 2> gpurun_out/ab_$ov.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$ov.json").read().strip().splitlines()[-1])
print("$VAR=$ov value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["rel_err"])
PY
done
env $VAR=1 timeout 300 python bench.py --workload cfg3_geom --steps 2 --warmup 2 --T 50 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 $VAR=1 fwd_ms', round(d['forward']['ms'],4), 'gcl_ms', d['roofline']['kernel_ms'])"
env $VAR=0 timeout 300 python bench.py --workload cfg3_geom --steps 2 --warmup 2 --T 50 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 $VAR=0 fwd_ms', round(d['forward']['ms'],4), 'gcl_ms', d['roofline']['kernel_ms'])"
