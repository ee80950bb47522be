#!/bin/bash
# Per-launch device times of the bench step (ncu, cold-cache & serialised: compare SHARES, not absolutes).
# usage: tools/launch_profile.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-front-end --only "$@" > gpurun_out/launches_${tag}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${tag}.csv > gpurun_out/launches_${tag}_summary.txt
head -24 gpurun_out/launches_${tag}_summary.txt
