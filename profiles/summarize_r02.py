#!/usr/bin/env python
"""Turn the artefacts of profiles/run_final_r02.sh (gpurun_out/final_r02/) into the tracked round-2 summaries under profiles/.
Usage: python profiles/summarize_r02.py [gpurun_out/final_r02] [tag]"""
import csv, glob, io, json, os, re, subprocess, sys

def edge_kind(kn):
    """'GCL' / 'COORD' from the last template argument of k_edge_v3<PROF, COORD> (ncu prints `<0, 1>` or `<(bool)0, (bool)1>`)."""
    m = re.search(r"k_edge_v3<([^>]*)>", kn)
    last = m.group(1).split(",")[-1].replace("(bool)", "").strip() if m else "0"
    return "COORD" if last in ("1", "true") else "GCL"

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/final_r02"
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
here = os.path.dirname(os.path.abspath(__file__))


def load(path):
    try:
        txt = open(path).read().strip().splitlines()
        return json.loads(txt[-1]) if txt else None
    except Exception:
        return None


# ---- configs table ---------------------------------------------------------------------------------------------
rows = []
order = ["bench_cfg2.json", "cfg_cfg2_zinc_L8.json", "cfg_cfg2_zinc_ragged.json", "cfg_cfg3_geom.json", "cfg_cfg4_pockets.json"] + \
        [f"cfg_cfg5_sweep_N{n}.json" for n in (32, 64, 128, 256, 512)]
for name in order:
    d = load(os.path.join(src, name))
    if not d:
        continue
    c, r, f = d["config"], d.get("roofline") or {}, d.get("forward") or {}
    e2e = d.get("e2e") or {}
    par = (d.get("parity") or {}).get("rel_err")
    rows.append("| {w} | {B} | {N} | {L} | {T} | {v:.1f} | {e} | {fm:.3f} | {k:.1f} | {a:.1f} | {fr:.4f} | {cf:.4f} | {hf:.4f} | {p} |".format(
        w=c["workload"], B=c["B"], N=c["N"], L=c["n_layers"], T=c["T"], v=d["value"],
        e=("%.1f" % e2e["value"]) if e2e else "-", fm=f.get("ms", 0), k=1e3 * r.get("kernel_ms", 0), a=r.get("achieved", 0),
        fr=r.get("frac", 0), cf=f.get("compute_frac", 0), hf=f.get("hbm_frac", 0), p=("%.1e" % par) if par is not None else "-"))
    if name == "bench_cfg2.json":
        json.dump(d, open(os.path.join(here, f"{tag}_bench_cfg2.json"), "w"))
ref = load(os.path.join(src, "bench_reference_arm.json"))
if ref:
    json.dump(ref, open(os.path.join(here, f"{tag}_bench_reference_arm.json"), "w"))
with open(os.path.join(here, f"{tag}_configs.md"), "w") as f:
    f.write(f"# BASELINE configs on 1x B200 ({tag}; `bash profiles/run_final_r02.sh` under gpurun, one pass)\n\n"
            "value = molecules/s of full T-step sampling, inputs resident in HBM, CUDA events over the timed steps, L2 flushed before every step;\n"
            "e2e = the same through `DDPM.sample_chain` from pinned host tensors incl. H2D/D2H; forward = device loop time / (T+1);\n"
            "GCL kernel = CUDA-event average of 20 isolated launches of the layer-0 GCL edge kernel; alg. TFLOP/s = (2H^2+10H) x edges per launch /\n"
            "kernel time; frac = alg. TFLOP/s / measured bf16 burst peak (MEASURED_PEAKS.json); compute_frac / hbm_frac = FLOPs_alg / Bytes_alg of a\n"
            "whole forward over the measured sustained peaks; parity = max rel. error of one Dynamics.forward at the timed shape vs the oracle.\n"
            "N <= 64 FC graphs run the third-generation edge kernels (k_edge_v3, GCL and COORD); N > 64 and the cut-off graphs of cfg4 the second-generation ones.\n\n"
            "| workload | B | N | L | T | molecules/s | e2e molecules/s | forward ms | GCL kernel us | GCL alg. TFLOP/s | frac | compute_frac | hbm_frac | parity |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(rows) + "\n")
    if ref:
        cb = ref["cpu_baseline"]
        f.write(f"\nReference arm (`bench.py --impl reference`, same box): {ref['value']:.4f} molecules/s -- kind `{cb['kind']}` ({cb['sample']}; {cb['cores']} torch threads); "
                f"oracle port beside it: {cb.get('port_value')}.\n")
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
        if "k_edge_v3" in r[ki]:
            name += " " + edge_kind(r[ki])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(here, f"{tag}_launches.md"), "w") as f:
        f.write(f"# Launch list of one forward ({tag}), cfg2_zinc B=256 N=40 L=6\n\n"
                "## ncu (cold-cache, serialised launches: compare SHARES, not absolutes)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 240 --csv python bench.py --steps 1 --warmup 1 --T 10 --no-e2e --no-cpu-baseline`\n\n"
                "| kernel | launches | avg us | total ms | share |\n|---|---|---|---|---|\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {us / n:.1f} | {us / 1e3:.3f} | {100 * us / tot:.1f}% |\n")
        lt = os.path.join(src, "live_kernel_times.txt")
        if os.path.isfile(lt):
            f.write("\n## live (DL_TIME_KERNELS=1: CUDA events around every launch of 23 eager `Dynamics.forward` calls, warm caches; each figure\n"
                    "includes ~3 us of launch gap that the CUDA-graph replay of the sampler does not pay -- per forward 12 GCL + 6 COORD + 13 node launches)\n\n```\n"
                    + open(lt).read() + "```\n")
        rc = os.path.join(src, "role_cycles.txt")
        if os.path.isfile(rc):
            f.write("\n## per-role cycle accounting (clock64 inside the kernels: DL_PROFILE_EDGE / DL_PROFILE_NODE)\n\n```\n" + open(rc).read() + "```\n")
        pl = os.path.join(src, "prof_live.txt")
        if os.path.isfile(pl):
            lines = open(pl).read().splitlines()
            coord = [l for l in lines if "COORD]" in l][:6]
            gcl = [l for l in lines if "GCL]" in l][6:12]
            f.write("\n## one eager forward, every edge launch profiled (DL_PROFILE_EDGE_LIVE: per-role cycles and the timeline of the first tile;\n"
                    "synchronised launches, so no overlap with the previous kernel; every timeline mark costs the marking warp a global round trip --\n"
                    "read it as an ordering with ~1 K cycles of overhead per mark on the same warp)\n\n```\n" + "\n".join(gcl + coord) + "\n```\n")
    print(open(os.path.join(here, f"{tag}_launches.md")).read())

# ---- full captures ---------------------------------------------------------------------------------------------
rep = os.path.join(src, "edge_node_full.ncu-rep")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(raw)))
    hdr, units = rws[0], rws[1]
    want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__block_size',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
    seen, out = set(), []
    traffic = {}
    for r in rws[2:]:
        d = dict(zip(hdr, r))
        kn = d['Kernel Name']
        key = edge_kind(kn) if "k_edge_v3" in kn else "node"
        if key in seen:
            continue
        seen.add(key)
        out.append(f"### {key}: `{kn[:110]}`\n")
        for w in want:
            if w in d:
                out.append(f"{w:75s} {d[w]} {units[hdr.index(w)]}")
        out.append("")
        try:
            def tobytes(name):
                v, u = float(d[name].replace(",", "")), units[hdr.index(name)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[key] = int(tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'))
        except Exception:
            pass
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rs = list(csv.reader(io.StringIO(srcp)))
    starts = [i for i, r in enumerate(rs) if r and r[0] == "Kernel Name"]
    done = set()
    for si, st in enumerate(starts):
        kn = rs[st][1]
        key = edge_kind(kn) if "k_edge_v3" in kn else "node"
        if key in done:
            continue
        done.add(key)
        blk = rs[st + 1: (starts[si + 1] if si + 1 < len(starts) else len(rs))]
        h, data = blk[0], blk[1:]
        ci = {x: i for i, x in enumerate(h)}
        stalls = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
        bounds = [0] + [i for i, r in enumerate(data) if 'USETMAXREG' in r[ci['Source']]] + [len(data)]
        names = ["prologue", "control (MMA issuer + table warps)", "producers", "epilogue"] if "k_edge_v3" in kn else ["whole kernel"] * 8
        out.append(f"### warp-stall samples by role, {key}\n")
        for bi in range(len(bounds) - 1):
            lo, hi = bounds[bi], bounds[bi + 1]
            tot = sum(int(r[ci['# Samples']]) for r in data[lo:hi])
            if tot == 0:
                continue
            agg = {s: sum(int(r[ci[s]]) for r in data[lo:hi]) for s in stalls}
            top = ", ".join(f"{s.replace('stall_', '')} {100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:7])
            out.append(f"* {names[min(bi, len(names) - 1)]} (SASS lines {lo}-{hi}, {tot} samples): {top}")
        out.append("")
    with open(os.path.join(here, f"{tag}_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full: GCL and COORD edge kernels (k_edge_v3) and the node kernel (k_node_tc2), {tag}\n\n"
                "Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_edge_v3|k_node_tc\" -s 14 -c 6 python bench.py --steps 1 --warmup 1 --T 4 --no-e2e --no-cpu-baseline`\n"
                "(first launch of each kind; ncu flushes caches between replays, so DRAM bytes are cold-cache figures: in the sampler the\n"
                "activations and tile tables of a forward stay in the 126 MB L2)\n\n```\n" + "\n".join(out) + "\n```\n")
    print("\n".join(out))
    if traffic:
        json.dump({"note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/{tag}_ncu_summary.md (ncu --set full, cold caches)",
                   "cfg2_zinc": {"edge_gcl": traffic.get("GCL"), "edge_coord": traffic.get("COORD"), "node": traffic.get("node")}},
                  open(os.path.join(here, "ncu_traffic.json"), "w"))

# ---- compute-sanitizer -------------------------------------------------------------------------------------------
sm = os.path.join(src, "sanitizer_memcheck.txt")
if os.path.isfile(sm):
    with open(os.path.join(here, f"{tag}_sanitizer.md"), "w") as f:
        f.write(f"# compute-sanitizer memcheck ({tag} build)\n\n`compute-sanitizer --tool memcheck python profiles/sanitize.py` on a B200 -- one small call of every native\n"
                "entry point: FC chain through the public `DDPM.sample_chain` (third-generation edge kernels, tile tables, device-side Philox noise,\n"
                "programmatic dependent launches inside the captured graph), InpaintingEDM chain, N=150 forward (second-generation kernels), cut-off graph forward,\n"
                "size classifier, bond orders, frame restore.\n\n```\n" + open(sm).read() + "```\n")
    print(open(os.path.join(here, f"{tag}_sanitizer.md")).read())

