#!/bin/bash
# k_p1_march: where the time goes -- march alone, boundary alone, per-CTA trace
mkdir -p gpurun_out
OUT=gpurun_out/r2ai_p1_march.txt
: > $OUT
for skip in 1 2; do
  for cfg in "1 8" "2 12" "2 16" "3 8"; do
    set -- $cfg
    LSB200_P1_MARCH_SKIP=$skip LSB200_P1_MARCH=$1 LSB200_P1_MARCH_CZ=$2 timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
  done
done
LSB200_P1_MARCH=2 LSB200_P1_MARCH_CZ=12 timeout 120 python scripts/cta_trace.py p1 117 > gpurun_out/r2ai_trace.txt 2>&1
LSB200_P1_MARCH_SKIP=1 LSB200_P1_MARCH=2 LSB200_P1_MARCH_CZ=12 timeout 120 python scripts/cta_trace.py p1 117 >> gpurun_out/r2ai_trace.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2ai_p1_march.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['env'].get('LSB200_P1_MARCH_SKIP'), d['env'].get('LSB200_P1_MARCH'), d['env'].get('LSB200_P1_MARCH_CZ'), round(d['matvec_us'], 2), d['plan'][:110])
PY
cat gpurun_out/r2ai_trace.txt
