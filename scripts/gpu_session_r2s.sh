#!/bin/bash
# Round 2, GPU session S (one B200): the plan chosen by timing (tune_fused) against the planner default
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fdi_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/s_pytest_fdi.txt 2>&1; tail -2 gpurun_out/s_pytest_fdi.txt
: > gpurun_out/s_matvec.jsonl
for n in 128 256 384 512; do
for env in "LSB200_X=0" "LSB200_AUTOTUNE=0"; do
  timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/s_matvec.jsonl 2>> gpurun_out/s_matvec.err
done
done
python - <<'PY'
import json
for l in open('gpurun_out/s_matvec.jsonl'):
    d=json.loads(l); print(d['n'], d['env'], round(d['matvec_us'],1), round(d['iteration_us'],1), round(d['matvec_gbs']), d['plan'])
PY
timeout 300 python scripts/prof_slab.py 1024 128 > gpurun_out/s_slab_1024.txt 2>&1; cat gpurun_out/s_slab_1024.txt
