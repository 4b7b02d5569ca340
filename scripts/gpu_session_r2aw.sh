#!/bin/bash
# k_p1_march on small tet meshes: chunks down to 4 planes, the boundary shell spread over the free CTA slots
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_p1_gpu.py tests/test_zz_plugin_device_gpu.py -x -q 2>&1 | tail -2
: > gpurun_out/aw_p1.txt
for n in 28 60; do timeout 100 python scripts/p1_experiment.py $n >> gpurun_out/aw_p1.txt 2>&1; done
LSB200_PERSIST=0 timeout 100 python scripts/p1_experiment.py 117 >> gpurun_out/aw_p1.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/aw_p1.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['cells'], d['env'], round(d['matvec_us'], 2), round(d['iteration_us'], 2), d['plan'][12:140])
PY
