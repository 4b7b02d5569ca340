#!/bin/bash
# config 3 (P1, 117^3 cells): persistent loop or launch-per-phase graph, k_p1_tiles or k_p1_march
mkdir -p gpurun_out
OUT=gpurun_out/r2am_c3.txt
: > $OUT
for persist in 1 0; do
  for march in 0 1 3; do
    LSB200_PERSIST=$persist LSB200_P1_MARCH=$march timeout 200 python scripts/small_systems.py p1 2>&1 | tail -1 >> $OUT
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2am_c3.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['env'], d['iterations'], round(d['solve_ms'], 1), round(d['us_per_iteration_solve'], 1), round(d['iteration_us'], 1), round(d['matvec_us'], 1), d['launches'])
PY
