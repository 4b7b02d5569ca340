#!/bin/bash
# Round 2, GPU session N (one B200): why the matvec of session M was slow (plan after the cost-model recalibration?)
set -u
mkdir -p gpurun_out
timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/n_trace_fdi_256.txt 2>&1; cat gpurun_out/n_trace_fdi_256.txt
: > gpurun_out/n_matvec.jsonl
for env in "LSB200_X=0" "LSB200_FUSED_SCZ=128" "LSB200_FUSED_SCZ=128 LSB200_FUSED_CZ=43" "LSB200_FUSED_SCZ=64" "LSB200_FUSED_SCZ=32"; do
  timeout 200 env $env python scripts/matvec_experiment.py 256 >> gpurun_out/n_matvec.jsonl 2>> gpurun_out/n_matvec.err
done
for env in "LSB200_X=0" "LSB200_FUSED_CZ=171 LSB200_FUSED_SCZ=128" "LSB200_FUSED_CZ=171"; do
  timeout 200 env $env python scripts/matvec_experiment.py 512 >> gpurun_out/n_matvec.jsonl 2>> gpurun_out/n_matvec.err
done
cut -c1-150 gpurun_out/n_matvec.jsonl
