#!/bin/bash
# Round 2, GPU session L (one B200): shared ranges after the restructure (one march call site, help only once every range
# is open), P1 tiles with the boundary share inside every CTA.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fdi_gpu.py tests/test_p1_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/l_pytest.txt 2>&1
tail -3 gpurun_out/l_pytest.txt
: > gpurun_out/l_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_SHARE=0" "LSB200_X=0" "LSB200_SHARE_EXTRA=0" "LSB200_SHARE_MIN=20" "LSB200_SHARE=0 LSB200_FUSED_CFG=3" "LSB200_FUSED_CFG=3"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/l_matvec.jsonl 2>> gpurun_out/l_matvec.err
  done
done
cut -c1-170 gpurun_out/l_matvec.jsonl
timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/l_trace_fdi_256.txt 2>&1; tail -3 gpurun_out/l_trace_fdi_256.txt
timeout 200 python scripts/cta_trace.py fdi 512 > gpurun_out/l_trace_fdi_512.txt 2>&1; tail -3 gpurun_out/l_trace_fdi_512.txt
for env in "LSB200_SHARE=0" "LSB200_X=0"; do
  echo "== $env" >> gpurun_out/l_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/l_slab_1024.txt 2>&1
done
cat gpurun_out/l_slab_1024.txt
: > gpurun_out/l_p1.jsonl
for env in "LSB200_P1_TILES=0" "LSB200_P1_TILE_CFG=0" "LSB200_P1_TILE_CFG=1" "LSB200_P1_TILE_CFG=2"; do
  timeout 200 env $env python scripts/p1_experiment.py 117 >> gpurun_out/l_p1.jsonl 2>> gpurun_out/l_p1.err
done
cut -c1-130 gpurun_out/l_p1.jsonl
timeout 200 python scripts/cta_trace.py p1 117 > gpurun_out/l_trace_p1_117.txt 2>&1; cat gpurun_out/l_trace_p1_117.txt
