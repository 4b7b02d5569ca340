#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/o_matvec.jsonl
for env in "LSB200_LIBRARY=/root/repo/loopstructural_b200/liblsb200_c8364db.so" "LSB200_X=0" "LSB200_FUSED_MODEL=12,1.45,2.0" "LSB200_FUSED_MODEL=12,0.35,0.45" "LSB200_FUSED_MODEL=16,1.45,2.0"  "LSB200_FUSED_MODEL=13,1.45,2.0"; do
  timeout 200 env $env python scripts/matvec_experiment.py 256 >> gpurun_out/o_matvec.jsonl 2>> gpurun_out/o_matvec.err
done
for env in "LSB200_LIBRARY=/root/repo/loopstructural_b200/liblsb200_c8364db.so" "LSB200_FUSED_MODEL=12,1.45,2.0"; do
  timeout 200 env $env python scripts/matvec_experiment.py 512 >> gpurun_out/o_matvec.jsonl 2>> gpurun_out/o_matvec.err
done
cut -c1-180 gpurun_out/o_matvec.jsonl; tail -3 gpurun_out/o_matvec.err
LSB200_FUSED_MODEL=12,1.45,2.0 timeout 200 python scripts/cta_trace.py fdi 256 | tail -6
