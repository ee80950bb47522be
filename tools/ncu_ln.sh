#!/bin/bash
# ncu full capture of the LayerNorm backward microbenchmark (tools/ln_bwd_bench.py), summarised on the box
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd -s 5 -c 1 -f -o /tmp/prof_ln python tools/ln_bwd_bench.py > /tmp/prof_ln.log 2>&1
python tools/ncu_top.py /tmp/prof_ln.ncu-rep > gpurun_out/ncu_ln_bwd.txt 2>&1
cat gpurun_out/ncu_ln_bwd.txt
