#!/bin/bash
# Round-end measurement pass (run under gpurun, ONE GPU): full GPU test suite, headline bench with e2e + CPU baseline,
# the other BASELINE configs, the ncu launch list of the bench command and one full ncu capture of the dominant kernel.
set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
for w in cfg2_zinc_L8 cfg2_zinc_ragged cfg3_geom; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_$w.json
done
timeout 500 python bench.py --workload cfg4_pockets --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_cfg4_pockets.json
for n in 32 64 128 256 512; do
  timeout 300 python bench.py --workload cfg5_sweep_N$n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > $O/cfg_cfg5_sweep_N$n.json
done
# launch list: kernels of the bench command (T shortened so that the capture window covers whole forwards)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 120 -c 240 --csv \
  --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --T 10 --no-e2e --no-cpu-baseline > $O/launches.log 2>&1
# one full capture of the GCL edge kernel (and the node kernel next to it)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_edge_tc|k_node_tc" -s 40 -c 3 \
  -o $O/edge_node_full python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
ls -la $O
