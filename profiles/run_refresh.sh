#!/bin/bash
# Short re-measurement after a kernel change: GPU tests, headline bench (e2e + CPU baseline), cfg3 / cfg4 / L8 lines.
set -x
O=gpurun_out/refresh; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
for w in cfg2_zinc_L8 cfg3_geom; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_$w.json
done
timeout 500 python bench.py --workload cfg4_pockets --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_cfg4_pockets.json
cat $O/pytest_gpu.txt
