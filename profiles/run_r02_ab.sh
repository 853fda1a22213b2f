#!/bin/bash
# A/B within one box: DL_TIME_KERNELS per-kernel live times + graph-replayed forward time from bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden and not T500 or config_shapes or odd or sweep" 2>&1 | tail -2
DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times"
python bench.py --steps 2 --warmup 2 --T 50 --no-e2e --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"], "parity", d["parity"]["rel_err"])
PY
