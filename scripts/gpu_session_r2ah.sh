#!/bin/bash
# k_p1_march: first run on the device -- parity tests with it as the default, then the sweep of configuration x chunk length
mkdir -p gpurun_out
OUT=gpurun_out/r2ah_p1_march.txt
: > $OUT
timeout 300 python -m pytest tests/test_p1_gpu.py tests/test_zz_iso_gpu.py -x -q > gpurun_out/r2ah_tests.txt 2>&1
tail -5 gpurun_out/r2ah_tests.txt
LSB200_P1_MARCH=0 timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
for m in 1 2 3; do
  for cz in 0 6 8 12 16 24; do
    LSB200_P1_MARCH=$m LSB200_P1_MARCH_CZ=$cz timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
  done
done
for cz in 0 12; do
  LSB200_P1_MARCH=4 LSB200_P1_MARCH_CZ=$cz timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2ah_p1_march.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['env'].get('LSB200_P1_MARCH'), d['env'].get('LSB200_P1_MARCH_CZ'), round(d['matvec_us'], 2), round(d['iteration_us'], 2), d['q_dot_probe'], d['p_dot_q'], d['plan'][:110])
PY
