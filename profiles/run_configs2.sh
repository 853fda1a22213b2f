#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py --workload cfg4_pockets --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/cfg_cfg4_pockets.json
for n in 256 512; do
  timeout 300 python bench.py --workload cfg5_sweep_N$n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/cfg_cfg5_sweep_N$n.json
done
