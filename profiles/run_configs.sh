#!/bin/bash
# Bench lines of the other BASELINE configs (parity-test cases, not the headline): run under gpurun.
set -x
mkdir -p gpurun_out
for w in cfg2_zinc_L8 cfg3_geom cfg2_zinc_ragged; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/cfg_$w.json
done
timeout 400 python bench.py --workload cfg4_pockets --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/cfg_cfg4_pockets.json
for n in 32 64 128 256 512; do
  timeout 300 python bench.py --workload cfg5_sweep_N$n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/cfg_cfg5_sweep_N$n.json
done
