#!/bin/bash
# Round 2, final single-GPU session: full device suite, default bench, per-kernel ncu evidence (launch list with DRAM bytes
# of a bench step and of the other kernels; --set full of the dominant kernels), P1 operator.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest.txt 2>&1
tail -4 gpurun_out/f_pytest.txt
timeout 900 python bench.py > gpurun_out/f_bench_256.json 2> gpurun_out/f_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline']['frac'], d['roofline']['traffic'], d['roofline_iteration']['frac'], d['e2e'], {k:v for k,v in d['eval'].items() if 'host' in k}, d.get('kernels'), d.get('other_configs'), d['config'].get('launch_plan'))
PY
timeout 300 python bench.py --n 512 --n-value 800000 --n-gradient 80000 --steps 2 --warmup 3 --no-experiments --no-cpu-baseline --no-fault --no-other-configs > gpurun_out/f_bench_512.json 2> gpurun_out/f_bench_512.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench_512.json'))
print('512:', {k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['config'].get('launch_plan'))
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/f_launches_bench_256.csv \
  python bench.py --steps 1 --warmup 1 --no-experiments --no-cpu-baseline --no-other-configs > gpurun_out/f_ncu_bench.log 2>&1
python scripts/ncu_summary.py gpurun_out/f_launches_bench_256.csv > gpurun_out/f_ncu_bench_256_summary.csv 2>&1; head -30 gpurun_out/f_ncu_bench_256_summary.csv
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_small.csv \
  python scripts/small_systems.py p1 claudius > gpurun_out/f_ncu_small.log 2>&1
python scripts/ncu_summary.py gpurun_out/f_launches_small.csv > gpurun_out/f_ncu_small_summary.csv 2>&1; head -12 gpurun_out/f_ncu_small_summary.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o /tmp/f_fused_256 -f python scripts/matvec_experiment.py 256 > gpurun_out/f_ncu_fused.log 2>&1
ncu -i /tmp/f_fused_256.ncu-rep --page raw --csv > gpurun_out/f_fused_256_raw.csv 2>/dev/null
ncu -i /tmp/f_fused_256.ncu-rep --page details --csv > gpurun_out/f_fused_256_details.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:k_p1_tiles -s 3 -c 1 -o /tmp/f_p1_tiles -f python scripts/p1_experiment.py 117 > gpurun_out/f_ncu_p1.log 2>&1
ncu -i /tmp/f_p1_tiles.ncu-rep --page details --csv > gpurun_out/f_p1_tiles_details.csv 2>/dev/null
ncu -i /tmp/f_p1_tiles.ncu-rep --page raw --csv > gpurun_out/f_p1_tiles_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:k_rows_normal -s 3 -c 1 -o /tmp/f_rows -f python scripts/prof_rows.py > gpurun_out/f_ncu_rows.log 2>&1
ncu -i /tmp/f_rows.ncu-rep --page details --csv > gpurun_out/f_rows_details.csv 2>/dev/null
: > gpurun_out/f_p1.jsonl
for env in "LSB200_P1_TILES=0" "LSB200_P1_TILE_CFG=0" "LSB200_P1_TILE_CFG=1" "LSB200_P1_TILE_CFG=2"; do
  timeout 200 env $env python scripts/p1_experiment.py 117 >> gpurun_out/f_p1.jsonl 2>> gpurun_out/f_p1.err
done
cut -c1-130 gpurun_out/f_p1.jsonl
ls -la gpurun_out | tail -20
