#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden and not T500 or config_shapes or odd or sweep" 2>&1 | tail -3
DL_PROFILE_NODE=1 python bench.py --steps 1 --warmup 1 --T 20 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/node_prof.json 2> gpurun_out/node_prof.err
grep "dl prof node" gpurun_out/node_prof.err | tail -4
python - <<PY
import json
d=json.loads(open("gpurun_out/node_prof.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"])
PY
DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times"
