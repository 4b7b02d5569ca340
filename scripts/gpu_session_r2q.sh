#!/bin/bash
# Round 2, GPU session Q (one B200): deeper strip ring (10 planes, 7 in flight), planner model sweep
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fdi_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/q_pytest_fdi.txt 2>&1; tail -2 gpurun_out/q_pytest_fdi.txt
: > gpurun_out/q_matvec.jsonl
for n in 256 512; do
for env in "LSB200_X=0" "LSB200_FUSED_MODEL=24,1.45,2.0" "LSB200_FUSED_MODEL=24,1.8,2.4" "LSB200_FUSED_MODEL=24,1.0,1.4" "LSB200_FUSED_MODEL=28,1.45,2.0" "LSB200_FUSED_MODEL=24,0.7,1.0" "LSB200_FUSED_MODEL=12,1.0,1.4"; do
  timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/q_matvec.jsonl 2>> gpurun_out/q_matvec.err
done
done
cut -c1-180 gpurun_out/q_matvec.jsonl; tail -3 gpurun_out/q_matvec.err
timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/q_trace_fdi_256.txt; tail -8 gpurun_out/q_trace_fdi_256.txt
for env in "LSB200_X=0" "LSB200_FUSED_MODEL=24,1.45,2.0"; do
  echo "== $env" >> gpurun_out/q_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/q_slab_1024.txt 2>&1
done
cat gpurun_out/q_slab_1024.txt
