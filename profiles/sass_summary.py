#!/usr/bin/env python
"""cuobjdump -sass of the shipped library -> per-kernel counts of the instructions that prove the Blackwell-native path
(B200_PROFILING.md "What proves a Blackwell-native kernel"). Usage: python profiles/sass_summary.py > profiles/sass_summary.md"""
import collections, os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
lib = os.path.join(os.path.dirname(here), "difflinker_b200", "libdifflinker_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "USETMAXREG", "ELECT", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.RCP",
        "F2FP", "HMMA", "LDGSTS", "STL", "LDL"]
kern = None
counts = collections.OrderedDict()
sizes = collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        continue
    sizes[kern] += 1
    for p in pats:
        if re.search(r"\b" + re.escape(p), line):
            counts[kern][p] += 1
dem = subprocess.run(["cu++filt"] + list(counts.keys()), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(counts.keys(), dem)) if len(dem) == len(counts) else {k: k for k in counts}
print("# SASS summary of libdifflinker_b200.so (sm_100a)\n")
print("`cuobjdump -sass difflinker_b200/libdifflinker_b200.so`, counted per kernel by `profiles/sass_summary.py` (static instruction counts).")
print("`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld / tcgen05.st (tensor memory), `UTCBAR` = tcgen05.commit, `UTMALDG` = cp.async.bulk.tensor")
print("(tiled TMA load), `UBLKCP` = cp.async.bulk (1-D TMA copy), `USETMAXREG` = setmaxnreg, `ELECT` = elect.sync, `SYNCS` = mbarrier ops,")
print("`FFMA2/FMUL2/FADD2` = packed fp32x2 arithmetic; `HMMA` (legacy mma.sync) and `LDGSTS` must be absent; `STL`/`LDL` = register spills.\n")
print("| kernel | SASS instrs | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
for k, c in counts.items():
    if not any(c[p] for p in pats[:7]) and sizes[k] < 400:
        continue
    short = re.sub(r"\(.*", "", names[k]).replace("dl::", "")
    targs = re.search(r"<[^>]*>", names[k])
    if targs and "k_edge" in short:
        short += " " + targs.group(0)
    print(f"| `{short[:60]}` | {sizes[k]} | " + " | ".join(str(c[p]) if c[p] else "" for p in pats) + " |")
