#!/bin/bash
# A/B within one box: the in-tree library (cur) against every experimental build profiles/ab/lib_*.so (git-ignored, built by hand
# from a patched copy of csrc/): forward parity tests on each, then graph-replayed forward + isolated GCL time, twice, alternating.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp difflinker_b200/libdifflinker_b200.so /tmp/lib_cur.so
one() {
  timeout 300 python bench.py --steps 3 --warmup 3 --T 100 --no-e2e --no-cpu-baseline "${@:2}" > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/abl.json").read().strip().splitlines()[-1])
print("$1 ${@:2} fwd_ms", round(d["forward"]["ms"],4), "gcl_ms", round(d["roofline"]["kernel_ms"],5), "parity", d["parity"]["rel_err"])
PY
}
for lib in profiles/ab/lib_*.so; do
  cp $lib difflinker_b200/libdifflinker_b200.so
  echo "== $lib: $(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'golden and not T500 or config_shapes or odd or sweep or nan' 2>&1 | tail -1)"
done
for rep in 1 2; do
  cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so; one cur
  for lib in profiles/ab/lib_*.so; do cp $lib difflinker_b200/libdifflinker_b200.so; one $(basename $lib .so); done
done
cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so; one cur --workload cfg3_geom
for lib in profiles/ab/lib_*.so; do cp $lib difflinker_b200/libdifflinker_b200.so; one $(basename $lib .so) --workload cfg3_geom; done
cp /tmp/lib_cur.so difflinker_b200/libdifflinker_b200.so
