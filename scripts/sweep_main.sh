#!/bin/bash
# usage: scripts/sweep_main.sh <n> <out>  -- CUDA-event time per launch of every stencil configuration
n=${1:-256}; out=${2:-gpurun_out/sweep_main.txt}
for m in -1 0 1 2 4; do echo -n "main=$m " >> $out; LSB200_STENCIL_MAIN=$m python scripts/prof_kernels.py $n 0 20 >> $out 2>&1; done
