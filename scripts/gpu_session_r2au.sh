#!/bin/bash
# k_p1_boundary: warps of one parity, short tap lists
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_p1_gpu.py -x -q 2>&1 | tail -2
: > gpurun_out/au_p1.txt
LSB200_PERSIST=0 timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/au_p1.txt 2>&1
timeout 100 python scripts/small_systems.py p1 2>&1 | tail -1 >> gpurun_out/au_p1.txt
LSB200_PERSIST=1 timeout 120 python scripts/p1_experiment.py 117 >> gpurun_out/au_p1.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/au_p1.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['env'], round(d['matvec_us'], 2), round(d['iteration_us'], 2), d.get('q_dot_probe'), d.get('solve_ms'), d.get('iterations'), d['plan'][12:80])
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_p1_boundary -s 5 -c 3 env LSB200_PERSIST=0 python scripts/p1_experiment.py 117 2>&1 | grep -i "duration" | head -3
