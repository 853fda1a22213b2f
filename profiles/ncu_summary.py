#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + warp-stall breakdown and hottest SASS lines.
Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_bytes.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in want:
        if w in d:
            print(f"{w:75s} {d[w]} {units[hdr.index(w)]}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = rows[starts[which] + 1: (starts[which + 1] if which + 1 < len(starts) else len(rows))]
print("kernel:", rows[starts[which]][1][:80])
h, data = blk[0], blk[1:]
ci = {x: i for i, x in enumerate(h)}
tot = sum(int(r[ci['# Samples']]) for r in data)
stalls = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
agg = {s: sum(int(r[ci[s]]) for r in data) for s in stalls}
print("samples", tot)
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"  {s:26s}{v:8d} {v / tot * 100:5.1f}%")
top = sorted(range(len(data)), key=lambda i: -int(data[i][ci['# Samples']]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for i in sorted(top):
    r = data[i]
    st = sorted(((s, int(r[ci[s]])) for s in stalls if int(r[ci[s]]) > 0), key=lambda x: -x[1])[:2]
    print(f"{i:5d} {r[ci['Source']].strip()[:64]:64s} {r[ci['# Samples']]:>6s} {st}")
