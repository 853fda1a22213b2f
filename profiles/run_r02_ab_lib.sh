#!/bin/bash
# A/B within one box of two builds: the in-tree library against profiles/ab/lib_new.so (built from an earlier commit)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not T500" 2>&1 | tail -3
cp difflinker_b200/libdifflinker_b200.so /tmp/lib_cur.so
one() {
  timeout 300 python bench.py --steps 3 --warmup 3 --T 100 --no-e2e --no-cpu-baseline "${@:2}" > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/abl.json").read().strip().splitlines()[-1])
print("$1 ${@:2} fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", round(d["roofline"]["kernel_ms"],5), "parity", d["parity"]["rel_err"])
PY
}
for rep in 1 2; do
  cp profiles/ab/lib_new.so difflinker_b200/libdifflinker_b200.so; one new
  cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so; one cur
done
cp profiles/ab/lib_new.so difflinker_b200/libdifflinker_b200.so; one new --workload cfg3_geom
cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so; one cur --workload cfg3_geom
for which in new cur; do
  if [ $which = new ]; then cp profiles/ab/lib_new.so difflinker_b200/libdifflinker_b200.so; else cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so; fi
  echo "== live kernel times, $which"
  DL_TIME_KERNELS=1 python profiles/time_kernels.py cfg2_zinc 20 2>&1 | grep "dl times" | grep "edge\|tiles"
done



