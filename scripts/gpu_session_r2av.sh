#!/bin/bash
# small grids: 64 x 8 tiles among the tuner's candidates
mkdir -p gpurun_out
timeout 200 python scripts/small_systems.py claudius fdi64 > gpurun_out/av_small.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/av_small.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['system'][:30], round(d['iteration_us'], 1), round(d['us_per_iteration_solve'], 1), round(d['matvec_us'], 1), d['iterations'], d.get('phases_us'), d['plan'][:150])
PY
timeout 400 python -m pytest tests/test_zz_claudius_gpu.py tests/test_fdi_gpu.py tests/test_zz_plugin_device_gpu.py -x -q -k "not full_size" 2>&1 | tail -3
