#!/bin/bash
# Round 2, GPU session U (one B200): the persistent CG loop of small systems -- device tests, iteration times against the
# launch-per-phase loop
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fdi_gpu.py tests/test_p1_gpu.py tests/test_zz_claudius_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/u_pytest.txt 2>&1
tail -6 gpurun_out/u_pytest.txt
: > gpurun_out/u_small.jsonl
for env in "LSB200_PERSIST=0" "LSB200_X=0"; do
  timeout 300 env $env python scripts/small_systems.py >> gpurun_out/u_small.jsonl 2>> gpurun_out/u_small.err
done
python - <<'PY'
import json
for l in open('gpurun_out/u_small.jsonl'):
    d=json.loads(l); print(d['system'], d['env'], 'its', d['iterations'], 'solve_ms %.2f' % d['solve_ms'], 'us/it %.1f' % d['us_per_iteration_solve'], 'iter_us %.1f' % d['iteration_us'], 'matvec %.1f' % d['matvec_us'], 'launches', d['launches'], 'chk %.9g' % d['checksum'])
PY
tail -3 gpurun_out/u_small.err
