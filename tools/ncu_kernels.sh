#!/bin/bash
# one full capture per kernel regex, summarised on the box (small text output only)
# usage: [BENCH_ARGS='--config cfg3'] tools/ncu_kernels.sh <tag> <skip> regex1 regex2 ...
tag=$1; skip=$2; shift 2
mkdir -p gpurun_out
for kre in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:${kre} -s ${skip} -c 1 -f -o /tmp/prof_${kre} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-front-end --only --no-graphs ${BENCH_ARGS} > /tmp/prof_${kre}.log 2>&1
  echo "=================== ${kre}" >> gpurun_out/ncu_${tag}.txt
  python tools/ncu_top.py /tmp/prof_${kre}.ncu-rep >> gpurun_out/ncu_${tag}.txt 2>&1
done
cat gpurun_out/ncu_${tag}.txt
