#!/bin/bash
# Round-2 measurement pass (run under gpurun, ONE GPU): full GPU test suite, headline bench with e2e + CPU baseline + reference
# arm, the other BASELINE configs, live per-kernel times, the ncu launch list of the bench command and full ncu captures of
# the GCL / COORD / node kernels. Artefacts land in gpurun_out/final_r02/; profiles/summarize_r02.py turns them into the
# tracked summaries under profiles/.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/final_r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in cfg2_zinc_L8 cfg2_zinc_ragged cfg3_geom; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_$w.json
done
timeout 500 python bench.py --workload cfg4_pockets --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/cfg_cfg4_pockets.json
for n in 32 64 128 256 512; do
  timeout 300 python bench.py --workload cfg5_sweep_N$n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > $O/cfg_cfg5_sweep_N$n.json
done
DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times" > $O/live_kernel_times.txt
DL_PROFILE_EDGE=1 DL_PROFILE_NODE=1 python bench.py --steps 1 --warmup 1 --T 20 --no-e2e --no-cpu-baseline 2>&1 >/dev/null | grep "dl prof" > $O/role_cycles.txt
DL_PROFILE_EDGE_LIVE=72 DL_TIME_KERNELS=0 python profiles/time_kernels.py cfg2_zinc 6 2>&1 | grep "dl prof v3" > $O/prof_live.txt
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize.py 2>&1 | grep -v "^=========\s*$" | tail -12 > $O/sanitizer_memcheck.txt
# launch list: kernels of the bench command (T shortened so that the capture window covers whole forwards)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 240 --csv \
  --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --T 10 --no-e2e --no-cpu-baseline > $O/launches.log 2>&1
# full captures: GCL (v3), COORD (v3), node kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_edge_v3|k_node_tc" -s 14 -c 6 \
  -o $O/edge_node_full python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
cat $O/pytest_gpu.txt
