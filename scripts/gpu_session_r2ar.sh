#!/bin/bash
# P1: the march and the boundary shell as two launches side by side
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_p1_gpu.py tests/test_zz_iso_gpu.py -x -q 2>&1 | tail -2
: > gpurun_out/ar_p1.txt
LSB200_PERSIST=0 timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/ar_p1.txt 2>&1
LSB200_PERSIST=0 LSB200_P1_SPLIT=0 timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/ar_p1.txt 2>&1
for cz in 8 10 17; do
  LSB200_PERSIST=0 LSB200_P1_MARCH_CZ=$cz timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/ar_p1.txt 2>&1
done
timeout 100 python scripts/small_systems.py p1 2>&1 | tail -1 >> gpurun_out/ar_p1.txt
LSB200_P1_SPLIT=0 timeout 100 python scripts/small_systems.py p1 2>&1 | tail -1 >> gpurun_out/ar_p1.txt
python - <<'PY'
import json
for l in open('gpurun_out/ar_p1.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['env'], round(d['matvec_us'], 2), round(d['iteration_us'], 2), d.get('q_dot_probe'), d.get('solve_ms'), d.get('iterations'), d['plan'][12:110])
PY
