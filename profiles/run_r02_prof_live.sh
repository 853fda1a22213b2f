#!/bin/bash
# per-role cycles and first-tile timeline of every edge launch of one eager forward (GCL and COORD)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DL_PROFILE_EDGE_LIVE=${1:-72} DL_TIME_KERNELS=0 python profiles/time_kernels.py ${2:-cfg2_zinc} 6 2>&1 | grep "dl prof v3" > gpurun_out/prof_live.txt
grep -c . gpurun_out/prof_live.txt
grep "COORD" gpurun_out/prof_live.txt | head -12
grep "GCL" gpurun_out/prof_live.txt | sed -n 7,12p
