#!/bin/bash
# instructions issued per launch of the edge kernels under the three wait modes (are failed polls a real share?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 1 2; do
  DL_WAIT_MODE=$m timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_edge_v3|k_node_tc2" -s 14 -c 4 --csv --log-file gpurun_out/inst_$m.csv python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  echo "== DL_WAIT_MODE=$m"; grep -v "^==" gpurun_out/inst_$m.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:40], r['Metric Name'], r['Metric Value'])"
done
