#!/bin/bash
# Round 2, GPU session V (one B200): even row pitch for every grid (TMA and 16-byte copies for odd nx), short-chunk plans for
# small grids, persistent CG loop -- full device suite and the small-system timings
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/v_pytest.txt 2>&1
tail -6 gpurun_out/v_pytest.txt
: > gpurun_out/v_small.jsonl
for env in "LSB200_PERSIST=0" "LSB200_X=0"; do
  timeout 300 env $env python scripts/small_systems.py >> gpurun_out/v_small.jsonl 2>> gpurun_out/v_small.err
done
python - <<'PY'
import json
for l in open('gpurun_out/v_small.jsonl'):
    d=json.loads(l); print(d['system'], d['env'], 'its', d['iterations'], 'solve_ms %.2f' % d['solve_ms'], 'us/it %.1f' % d['us_per_iteration_solve'], 'iter_us %.1f' % d['iteration_us'], 'matvec %.1f' % d['matvec_us'], 'launches', d['launches'], 'chk %.9g' % d['checksum'], d['plan'][:150])
PY
tail -3 gpurun_out/v_small.err
timeout 100 python scripts/matvec_experiment.py 256 | cut -c1-200
