#!/bin/bash
# persistent loop: CTA 0's phase clock on the three small systems
mkdir -p gpurun_out
timeout 200 python scripts/small_systems.py claudius fdi64 > gpurun_out/at_small.txt 2>&1
LSB200_PERSIST=1 timeout 200 python scripts/small_systems.py p1 >> gpurun_out/at_small.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/at_small.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['system'][:30], round(d['iteration_us'], 1), round(d['us_per_iteration_solve'], 1), d.get('phases_us'))
PY
