#!/bin/bash
# k_p1_march: 2x2 patches against 2x1 patches (twice the marching warps), each twice
mkdir -p gpurun_out
OUT=gpurun_out/r2ak_p1_march.txt
: > $OUT
timeout 300 python -m pytest tests/test_p1_gpu.py -x -q > gpurun_out/r2ak_tests.txt 2>&1
tail -3 gpurun_out/r2ak_tests.txt
for rep in 1 2; do
  for m in 0 1 3 5 6 7; do
    LSB200_P1_MARCH=$m timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
  done
done
for m in 5 6 7; do
  LSB200_P1_MARCH=$m LSB200_P1_MARCH_CZ=9 timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
done
LSB200_P1_MARCH=6 timeout 120 python scripts/cta_trace.py p1 117 > gpurun_out/r2ak_trace.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2ak_p1_march.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['env'].get('LSB200_P1_MARCH'), d['env'].get('LSB200_P1_MARCH_CZ'), round(d['matvec_us'], 2), round(d['iteration_us'], 2), d['q_dot_probe'], d['p_dot_q'], d['plan'][12:130])
PY
head -4 gpurun_out/r2ak_trace.txt; tail -2 gpurun_out/r2ak_trace.txt
