#!/bin/bash
# per-kernel launch list of one forward (ncu, cold-cache serialised: compare shares) + node-kernel phase marks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DL_PROFILE_NODE=1 python bench.py --steps 1 --warmup 1 --T 20 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/node_prof.json 2> gpurun_out/node_prof.err
grep "dl prof node" gpurun_out/node_prof.err | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 240 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --T 10 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    name = r[ki].split("(")[0][-40:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:42s} {n:4d} launches  avg {t/n/1000:8.2f} us  share {t/tot*100:5.1f}%")
PY
