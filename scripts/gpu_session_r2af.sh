#!/bin/bash
# Round 2, GPU session AF (one B200): graph kept across solves; L2 persistence of p (experiment)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fdi_gpu.py tests/test_zz_claudius_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/af_pytest.txt 2>&1; tail -2 gpurun_out/af_pytest.txt
: > gpurun_out/af_matvec.jsonl
for env in "LSB200_X=0" "LSB200_L2_PERSIST=32" "LSB200_L2_PERSIST=64" "LSB200_L2_PERSIST=96"; do
  timeout 200 env $env python scripts/matvec_experiment.py 256 >> gpurun_out/af_matvec.jsonl 2>> gpurun_out/af_matvec.err
done
python - <<'PY'
import json
for l in open('gpurun_out/af_matvec.jsonl'):
    d=json.loads(l); print(d['n'], d['env'], round(d['matvec_us'],1), round(d['iteration_us'],1))
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-experiments --no-cpu-baseline --no-fault --no-other-configs > gpurun_out/af_bench_256.json 2> gpurun_out/af_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/af_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value','solve_us_per_iteration')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['e2e']['ms_per_step'])
PY
