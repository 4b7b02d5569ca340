#!/bin/bash
# k_p1_march with boundary warps: parity tests, sweep of configuration x chunk length (operator and persistent CG loop), traces
mkdir -p gpurun_out
OUT=gpurun_out/r2aj_p1_march.txt
: > $OUT
timeout 300 python -m pytest tests/test_p1_gpu.py tests/test_zz_iso_gpu.py -x -q > gpurun_out/r2aj_tests.txt 2>&1
tail -5 gpurun_out/r2aj_tests.txt
LSB200_P1_MARCH=0 timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
for m in 1 2 3 4; do
  for cz in 0 8 16 20; do
    LSB200_P1_MARCH=$m LSB200_P1_MARCH_CZ=$cz timeout 120 python scripts/p1_experiment.py 117 >> $OUT 2>&1
  done
done
for m in 2 3; do
  LSB200_P1_MARCH=$m timeout 120 python scripts/cta_trace.py p1 117 >> gpurun_out/r2aj_trace.txt 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2aj_p1_march.txt'):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['env'].get('LSB200_P1_MARCH'), d['env'].get('LSB200_P1_MARCH_CZ'), round(d['matvec_us'], 2), round(d['iteration_us'], 2), d['q_dot_probe'], d['p_dot_q'], d['plan'][12:130])
PY
cat gpurun_out/r2aj_trace.txt
# one ncu capture of the default configuration (source-level stall reasons)
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_p1_march -s 5 -c 1 -o gpurun_out/r2aj_p1_march python scripts/p1_experiment.py 117 > gpurun_out/r2aj_ncu.log 2>&1
ncu -i gpurun_out/r2aj_p1_march.ncu-rep --page details --csv > gpurun_out/r2aj_ncu_details.csv 2>/dev/null
ncu -i gpurun_out/r2aj_p1_march.ncu-rep --page source --csv > gpurun_out/r2aj_ncu_source.csv 2>/dev/null
ls -la gpurun_out/r2aj_* | head
rm -f gpurun_out/r2aj_p1_march.ncu-rep
