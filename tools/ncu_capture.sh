#!/bin/bash
# Full ncu capture of a few launches of one kernel; exports the raw metrics page as CSV and keeps the .ncu-rep only
# when it is small enough to travel back (gpurun_out is capped at 64 MiB).
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> [bench args...]
tag=$1; kre=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${kre} -s ${skip} -c ${cnt} -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-front-end --only "$@" > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/prof_${tag}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/prof_${tag}.ncu-rep; fi
ls -la gpurun_out/ | grep prof_${tag}
