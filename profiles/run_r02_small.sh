#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden and not T500 or config_shapes" 2>&1 | tail -2
DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_finish|k_prep|k_tiles_d" -s 6 -c 3 -o gpurun_out/small_full python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log | cut -c1-80
