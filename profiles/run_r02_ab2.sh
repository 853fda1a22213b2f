#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden and not T500 or config_shapes or odd or sweep or nan or equivariance" 2>&1 | tail -3
for v in "DL_X=1" "DL_NODE_V1=1"; do
  echo "== $v"
  env $v DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times" | grep node
  env $v DL_PROFILE_NODE=1 python bench.py --steps 2 --warmup 2 --T 50 --no-e2e --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  grep "dl prof node" gpurun_out/ab.err | tail -2
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["rel_err"])
PY
done
