#!/bin/bash
# quick round-2 iteration: TS-mode probe, forward parity, then the GCL kernel timed in both generations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "selftest or golden or config_shapes or odd or simt_and" 2>&1 | tail -15 > gpurun_out/quick_tests.log
for v in 1 0; do
  DL_EDGE_V3=$v DL_PROFILE_EDGE=1 python bench.py --steps 1 --warmup 1 --T 20 --no-e2e --no-cpu-baseline > gpurun_out/quick_bench_v3_$v.json 2> gpurun_out/quick_bench_v3_$v.err
done
tail -5 gpurun_out/quick_tests.log
for v in 1 0; do grep "dl prof" gpurun_out/quick_bench_v3_$v.err | tail -6; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/quick_bench_v3_$v.json").read().strip().splitlines()[-1])
    print("v3=$v", "value", round(d["value"],1), "fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", d["roofline"]["kernel_ms"])
except Exception as ex:
    print("v3=$v failed", ex); print(open("gpurun_out/quick_bench_v3_$v.err").read()[-1500:])
PY
done
