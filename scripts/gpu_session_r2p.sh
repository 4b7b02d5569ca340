#!/bin/bash
# Round 2, GPU session P (one B200): face cost of the planner's model, 256^3 / 512^3 / slab
set -u
mkdir -p gpurun_out
: > gpurun_out/p_matvec.jsonl
for n in 256 512; do
for env in "LSB200_X=0" "LSB200_FUSED_MODEL=16,1.45,2.0" "LSB200_FUSED_MODEL=20,1.45,2.0" "LSB200_FUSED_MODEL=24,1.45,2.0" "LSB200_FUSED_MODEL=20,1.2,1.6" "LSB200_FUSED_MODEL=20,1.8,2.4"; do
  timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/p_matvec.jsonl 2>> gpurun_out/p_matvec.err
done
done
cut -c1-180 gpurun_out/p_matvec.jsonl; tail -3 gpurun_out/p_matvec.err
for env in "LSB200_X=0" "LSB200_FUSED_MODEL=20,1.45,2.0"; do
  echo "== $env" >> gpurun_out/p_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/p_slab_1024.txt 2>&1
done
cat gpurun_out/p_slab_1024.txt
LSB200_FUSED_MODEL=20,1.45,2.0 timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/p_trace_fdi_256_face20.txt; tail -8 gpurun_out/p_trace_fdi_256_face20.txt
