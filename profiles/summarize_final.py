#!/usr/bin/env python
"""Turn the artefacts of profiles/run_final.sh (gpurun_out/final/) into the tracked summaries under profiles/.
Usage: python profiles/summarize_final.py [gpurun_out/final] [tag]"""
import csv, glob, io, json, os, re, subprocess, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/final"
tag = sys.argv[2] if len(sys.argv) > 2 else "r01_final"
here = os.path.dirname(os.path.abspath(__file__))


def load(path):
    try:
        txt = open(path).read().strip().splitlines()
        return json.loads(txt[-1]) if txt else None
    except Exception:
        return None


# ---- configs table ---------------------------------------------------------------------------------------------
rows = []
for path in [os.path.join(src, "bench_cfg2.json")] + sorted(glob.glob(os.path.join(src, "cfg_*.json"))):
    d = load(path)
    if not d:
        continue
    c, r, f = d["config"], d.get("roofline") or {}, d.get("forward") or {}
    e2e = d.get("e2e") or {}
    rows.append("| {w} | {B} | {N} | {L} | {T} | {v:.1f} | {e} | {fm:.3f} | {k:.1f} | {a:.1f} | {fr:.4f} | {cf:.4f} | {hf:.4f} | {edges} |".format(
        w=c["workload"], B=c["B"], N=c["N"], L=c["n_layers"], T=c["T"], v=d["value"],
        e=("%.1f" % e2e["value"]) if e2e else "-", fm=f.get("ms", 0), k=1e3 * r.get("kernel_ms", 0), a=r.get("achieved", 0),
        fr=r.get("frac", 0), cf=f.get("compute_frac", 0), hf=f.get("hbm_frac", 0), edges=c.get("edges_per_launch", "-")))
    if path.endswith("bench_cfg2.json"):
        json.dump(d, open(os.path.join(here, f"{tag}_bench_cfg2.json"), "w"))
with open(os.path.join(here, f"{tag}_configs.md"), "w") as f:
    f.write(f"# BASELINE configs on 1x B200 ({tag}; `bash profiles/run_final.sh` under gpurun)\n\n"
            "value = molecules/s of full T-step sampling, inputs resident in HBM, CUDA events over the timed steps; e2e = the same through\n"
            "`DDPM.sample_chain` from pinned host tensors incl. H2D/D2H; forward = device loop time / (T+1); GCL kernel = CUDA-event average of\n"
            "20 isolated launches of the layer-0 GCL edge kernel; alg. TFLOP/s = (2H^2+10H) x edges per launch / kernel time (cut-off graphs:\n"
            "the edges the kernel walked, `dl_cut_graph_stats`); frac = alg. TFLOP/s / measured bf16 peak (MEASURED_PEAKS.json);\n"
            "compute_frac / hbm_frac = FLOPs_alg / Bytes_alg of a whole forward over the measured sustained peaks.\n\n"
            "| workload | B | N | L | T | molecules/s | e2e molecules/s | forward ms | GCL kernel us | GCL alg. TFLOP/s | frac | compute_frac | hbm_frac | edges/launch |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(rows) + "\n")
print("\n".join(rows))

# ---- launch list -----------------------------------------------------------------------------------------------
lp = os.path.join(src, "launches.csv")
if os.path.isfile(lp):
    txt = open(lp).read()
    start = txt.find('"ID"')
    rd = list(csv.reader(io.StringIO(txt[start:])))
    hdr = rd[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rd[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[ki]).replace("dl::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(here, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}), cfg2_zinc B=256 N=40 L=6\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 120 -c 240 --csv python bench.py --steps 1 --warmup 1 --T 10 --no-e2e --no-cpu-baseline`\n"
                "(cold-cache, serialised launches: compare SHARES with the live CUDA-event figures of bench.py, not absolutes).\n\n"
                "| kernel | launches | avg us | total ms | share |\n|---|---|---|---|---|\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {us / n:.1f} | {us / 1e3:.3f} | {100 * us / tot:.1f}% |\n")
    print(open(os.path.join(here, f"{tag}_launches.md")).read())

# ---- full capture ----------------------------------------------------------------------------------------------
rep = os.path.join(src, "edge_node_full.ncu-rep")
if os.path.isfile(rep):
    out = subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), rep, "0", "30"], capture_output=True, text=True)
    with open(os.path.join(here, f"{tag}_edge_tc_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full, GCL edge kernel (+ node kernel), {tag}\n\n"
                "Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_edge_tc|k_node_tc\" -s 40 -c 3 python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline`\n\n```\n"
                + out.stdout + "\n```\n")
    print(out.stdout[-3000:], out.stderr[-500:])
