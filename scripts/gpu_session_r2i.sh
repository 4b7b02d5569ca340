#!/bin/bash
# Round 2, GPU session I (one B200): ncu --set full of k_fdi_fused at 256^3 (TMA / cp.async) and 512^3 (TMA / cp.async)
set -u
mkdir -p gpurun_out
for n in 256 512; do
  for tma in 1 0; do
    LSB200_TMA=$tma timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 \
      -o gpurun_out/i_fused_${n}_tma${tma} -f python scripts/matvec_experiment.py $n > gpurun_out/i_ncu_${n}_${tma}.log 2>&1
    tail -1 gpurun_out/i_ncu_${n}_${tma}.log | cut -c1-200
  done
done
ls -la gpurun_out/*.ncu-rep
