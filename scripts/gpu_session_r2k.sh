#!/bin/bash
# Round 2, GPU session K (one B200): shared ranges of k_fdi_fused (two-ended marches) -- device tests with the default
# and with every range helped (LSB200_SHARE_MIN=1), timelines, matvec times against the static plan.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fdi_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/k_pytest_fdi.txt 2>&1
tail -3 gpurun_out/k_pytest_fdi.txt
LSB200_SHARE_MIN=1 timeout 600 python -m pytest tests/test_fdi_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/k_pytest_fdi_min1.txt 2>&1
tail -3 gpurun_out/k_pytest_fdi_min1.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/matvec_experiment.py 96 > gpurun_out/k_memcheck.txt 2>&1; tail -4 gpurun_out/k_memcheck.txt | cut -c1-200
: > gpurun_out/k_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_SHARE=0" "LSB200_X=0" "LSB200_SHARE_MIN=6" "LSB200_SHARE_MIN=24" "LSB200_FUSED_CFG=3" "LSB200_TMA=0"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/k_matvec.jsonl 2>> gpurun_out/k_matvec.err
  done
done
cut -c1-170 gpurun_out/k_matvec.jsonl
timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/k_trace_fdi_256.txt 2>&1; cat gpurun_out/k_trace_fdi_256.txt
timeout 200 python scripts/cta_trace.py fdi 512 > gpurun_out/k_trace_fdi_512.txt 2>&1; tail -4 gpurun_out/k_trace_fdi_512.txt
for env in "LSB200_SHARE=0" "LSB200_X=0"; do
  echo "== $env" >> gpurun_out/k_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/k_slab_1024.txt 2>&1
done
cat gpurun_out/k_slab_1024.txt
