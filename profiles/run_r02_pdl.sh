#!/bin/bash
# A/B within one box: programmatic dependent launch between the kernels of a forward on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not T500" 2>&1 | tail -3
for ov in 0 1 0 1; do
  DL_CHAIN_OVERLAP=$ov timeout 300 python bench.py --steps 3 --warmup 3 --T 100 --no-e2e --no-cpu-baseline > gpurun_out/pdl_$ov.json 2> gpurun_out/pdl_$ov.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pdl_$ov.json").read().strip().splitlines()[-1])
print("overlap=$ov value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["rel_err"])
PY
done
