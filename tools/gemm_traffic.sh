#!/bin/bash
# DRAM traffic of every tensor-core GEMM launch of ONE bench step (ncu, dram__bytes_{read,write}.sum + duration).
# usage: tools/gemm_traffic.sh <tag> <gemm launches per step> [bench args...]
tag=$1; per_step=$2; shift 2
mkdir -p gpurun_out
skip=$((3 * per_step))
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 \
    -s ${skip} -c ${per_step} --csv --log-file gpurun_out/gemm_traffic_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-front-end --only --no-graphs "$@" > gpurun_out/gemm_traffic_${tag}.log 2>&1
python tools/summarize_traffic.py gpurun_out/gemm_traffic_${tag}.csv > gpurun_out/gemm_traffic_${tag}.json
cat gpurun_out/gemm_traffic_${tag}.json
