#!/bin/bash
# Round 2, GPU session I (one B200): ncu --set full of k_fdi_fused at 256^3 (TMA / cp.async) and 512^3 (TMA / cp.async);
# the reports are exported to CSV on the box (gpurun_out/ carries at most 64 MiB back)
set -u
mkdir -p gpurun_out
for n in 256 512; do
  for tma in 1 0; do
    LSB200_TMA=$tma timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 \
      -o /tmp/i_fused_${n}_tma${tma} -f python scripts/matvec_experiment.py $n > gpurun_out/i_ncu_${n}_${tma}.log 2>&1
    ncu -i /tmp/i_fused_${n}_tma${tma}.ncu-rep --page raw --csv > gpurun_out/i_fused_${n}_tma${tma}_raw.csv 2>/dev/null
    ncu -i /tmp/i_fused_${n}_tma${tma}.ncu-rep --page details --csv > gpurun_out/i_fused_${n}_tma${tma}_details.csv 2>/dev/null
    ncu -i /tmp/i_fused_${n}_tma${tma}.ncu-rep --page source --csv > gpurun_out/i_fused_${n}_tma${tma}_source.csv 2>/dev/null
  done
done
cp /tmp/i_fused_256_tma1.ncu-rep gpurun_out/
ls -la gpurun_out/
