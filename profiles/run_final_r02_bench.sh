#!/bin/bash
# Bench-only part of the round-2 measurement pass (same output directory; see run_final_r02.sh)
# Round-2 measurement pass (run under gpurun, ONE GPU): full GPU test suite, headline bench with e2e + CPU baseline + reference
# arm, the other BASELINE configs, live per-kernel times, the ncu launch list of the bench command and full ncu captures of
# the GCL / COORD / node kernels. Artefacts land in gpurun_out/final_r02/; profiles/summarize_r02.py turns them into the
# tracked summaries under profiles/.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/final_r02; mkdir -p $O
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in cfg2_zinc_L8 cfg2_zinc_ragged cfg3_geom; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_$w.json
done
timeout 500 python bench.py --workload cfg4_pockets --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_cfg4_pockets.json
for n in 32 64 128 256 512; do
  timeout 300 python bench.py --workload cfg5_sweep_N$n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > $O/cfg_cfg5_sweep_N$n.json
done
ls -la $O | head -30
