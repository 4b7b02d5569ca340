#!/bin/bash
# Round 2, GPU session AE (one B200): 64 x 8 box tiles (5 CTAs per SM) among the tuner's candidates
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fdi_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/ae_pytest_fdi.txt 2>&1; tail -2 gpurun_out/ae_pytest_fdi.txt
: > gpurun_out/ae_matvec.jsonl
for n in 128 256 384 512; do
for env in "LSB200_X=0" "LSB200_FUSED_CFG=2"; do
  timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/ae_matvec.jsonl 2>> gpurun_out/ae_matvec.err
done
done
python - <<'PY'
import json
for l in open('gpurun_out/ae_matvec.jsonl'):
    d=json.loads(l); print(d['n'], d['env'], round(d['matvec_us'],1), round(d['iteration_us'],1), round(d['matvec_gbs']), d['plan'])
PY
timeout 300 python scripts/prof_slab.py 1024 128 > gpurun_out/ae_slab_1024.txt 2>&1; cat gpurun_out/ae_slab_1024.txt
timeout 100 python scripts/p1_experiment.py 117 | cut -c1-150
