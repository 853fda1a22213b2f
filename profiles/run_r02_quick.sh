#!/bin/bash
# quick round-2 iteration: forward parity, then the GCL kernel timed in variants given as "NAME:ENV=VAL,ENV=VAL" arguments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "selftest or golden or config_shapes or odd or simt_and" 2>&1 | tail -15 > gpurun_out/quick_tests.log
tail -3 gpurun_out/quick_tests.log
fi
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  envs="${envs//,/ }"
  env $envs DL_PROFILE_EDGE=1 python bench.py --steps 1 --warmup 1 --T 20 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/quick_$name.json 2> gpurun_out/quick_$name.err
  grep "dl prof" gpurun_out/quick_$name.err | tail -6
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/quick_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"])
except Exception as ex:
    print("$name failed", ex); print(open("gpurun_out/quick_$name.err").read()[-1500:])
PY
done
