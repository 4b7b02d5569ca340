#!/bin/bash
# Round 2, last single-GPU session: full device suite, default bench, ncu evidence of the P1 operator as shipped
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f2_pytest.txt 2>&1
tail -3 gpurun_out/f2_pytest.txt
timeout 900 python bench.py > gpurun_out/f2_bench_256.json 2> gpurun_out/f2_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f2_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['e2e'], d.get('other_configs'))
PY
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/f2_launches_c3.csv \
  python scripts/small_systems.py p1 > gpurun_out/f2_ncu_c3.log 2>&1
python scripts/ncu_summary.py gpurun_out/f2_launches_c3.csv > gpurun_out/f2_c3_summary.csv 2>&1; head -8 gpurun_out/f2_c3_summary.csv
timeout 200 ncu --set full --clock-control none -k regex:k_p1_march -s 5 -c 1 -o /tmp/f2_march -f env LSB200_PERSIST=0 python scripts/p1_experiment.py 117 > gpurun_out/f2_ncu_march.log 2>&1
ncu -i /tmp/f2_march.ncu-rep --page details --csv > gpurun_out/f2_ncu_full_k_p1_march_details.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:k_p1_boundary -s 5 -c 1 -o /tmp/f2_bnd -f env LSB200_PERSIST=0 python scripts/p1_experiment.py 117 > gpurun_out/f2_ncu_bnd.log 2>&1
ncu -i /tmp/f2_bnd.ncu-rep --page details --csv > gpurun_out/f2_ncu_full_k_p1_boundary_details.csv 2>/dev/null
ls -la gpurun_out/f2_* | head -12
