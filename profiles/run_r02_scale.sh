#!/bin/bash
# multi-GPU lines: usage run_r02_scale.sh N [T-override]
cd "$(dirname "$0")/.."
N=${1:-8}; TARG=""; [ -n "$2" ] && TARG="--T $2"
OUT=gpurun_out/scale_r02; mkdir -p $OUT
run() { # name, args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" $TARG > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
  echo "$name n=$N rc=$?: $(tail -1 $OUT/${name}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), d['unit'], 'scaling', d['scaling'], 'ms/step', round(d['ms_per_step'],1), 'e2e', d.get('e2e') and round(d['e2e']['value'],1))
except Exception as ex: print('no json', ex)")"
}
run cfg2_weak --workload cfg2_zinc --steps 3 --warmup 3 --no-cpu-baseline
run cfg3_strong --workload cfg3_geom --scaling strong --steps 3 --warmup 3 --no-cpu-baseline
run cfg3_weak --workload cfg3_geom --steps 2 --warmup 3 --no-cpu-baseline --no-e2e
run cfg4_strong --workload cfg4_pockets --scaling strong --steps 2 --warmup 3 --no-cpu-baseline
run cfg4_weak --workload cfg4_pockets --steps 2 --warmup 3 --no-cpu-baseline --no-e2e
for f in $OUT/*_n$N.err; do grep -il "error\|Traceback" $f; done | head
