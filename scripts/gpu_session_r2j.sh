#!/bin/bash
# Round 2, GPU session J (one B200): k_p1_tiles (device tests, tile shapes against the generic kernel at config-3 size),
# per-CTA timelines of k_fdi_fused (256^3, 512^3) and k_p1_tiles.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_p1_gpu.py tests/test_zz_e_p1_slab_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/j_pytest_p1.txt 2>&1
tail -3 gpurun_out/j_pytest_p1.txt
: > gpurun_out/j_p1.jsonl
for env in "LSB200_P1_TILES=0" "LSB200_P1_TILE_CFG=0" "LSB200_P1_TILE_CFG=1" "LSB200_P1_TILE_CFG=2"; do
  timeout 200 env $env python scripts/p1_experiment.py 117 >> gpurun_out/j_p1.jsonl 2>> gpurun_out/j_p1.err
done
cut -c1-200 gpurun_out/j_p1.jsonl
timeout 200 python scripts/cta_trace.py fdi 256 > gpurun_out/j_trace_fdi_256.txt 2>&1; cat gpurun_out/j_trace_fdi_256.txt
timeout 200 python scripts/cta_trace.py fdi 512 > gpurun_out/j_trace_fdi_512.txt 2>&1; cat gpurun_out/j_trace_fdi_512.txt
timeout 200 python scripts/cta_trace.py p1 117 > gpurun_out/j_trace_p1_117.txt 2>&1; cat gpurun_out/j_trace_p1_117.txt
timeout 200 python scripts/matvec_experiment.py 256 | cut -c1-200
