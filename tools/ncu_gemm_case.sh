#!/bin/bash
# full ncu capture of one row of tools/gemm_sweep.py (4th launch), summarised on the box
# usage: tools/ncu_gemm_case.sh <tag> <row> [<row> ...]
tag=$1; shift
mkdir -p gpurun_out
for row in "$@"; do
  SWEEP_ONE=$row ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o /tmp/prof_case${row} \
      python tools/gemm_sweep.py > /tmp/prof_case${row}.log 2>&1
  echo "=================== gemm_sweep row ${row}: $(tail -1 /tmp/prof_case${row}.log)" >> gpurun_out/ncu_${tag}.txt
  python tools/ncu_top.py /tmp/prof_case${row}.ncu-rep >> gpurun_out/ncu_${tag}.txt 2>&1
done
cat gpurun_out/ncu_${tag}.txt
