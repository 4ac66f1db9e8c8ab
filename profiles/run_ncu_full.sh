#!/bin/bash
# ncu --set full captures of selected launches of one bench step (C2, 1 GPU); the reports are exported to CSV on the box
# (raw metrics + source-level stall samples) and the .ncu-rep files are dropped when they would exceed the copy-back limit.
# usage: run_ncu_full.sh <tag> <kernel-regex> <skip> <count>
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=$3; CNT=$4
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_$TAG.log 2>&1; echo "$TAG rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>/dev/null
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 12000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
du -sh gpurun_out
