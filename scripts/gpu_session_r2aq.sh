#!/bin/bash
# k_p1_march: boundary nodes two at a time
mkdir -p gpurun_out
: > gpurun_out/aq_p1.txt
for m in 3 1 2; do
  LSB200_P1_MARCH=$m timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/aq_p1.txt 2>&1
done
LSB200_P1_MARCH=3 LSB200_PERSIST=1 timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/aq_p1.txt 2>&1
timeout 100 python scripts/small_systems.py p1 2>&1 | tail -1 >> gpurun_out/aq_p1.txt
LSB200_PERSIST=1 timeout 100 python scripts/small_systems.py p1 2>&1 | tail -1 >> gpurun_out/aq_p1.txt
python - <<'PY'
import json
for l in open('gpurun_out/aq_p1.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['env'], round(d['matvec_us'], 2), round(d['iteration_us'], 2), d.get('q_dot_probe'), d.get('solve_ms'), d.get('iterations'), d['plan'][12:100])
PY
timeout 300 python -m pytest tests/test_p1_gpu.py -x -q 2>&1 | tail -2
