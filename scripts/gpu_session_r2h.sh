#!/bin/bash
# Round 2, GPU session H (one B200), first call after the container was re-created: full device suite, the fused /
# TMA / three-launch matvec plans side by side, default bench, ncu launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/h_pytest.txt 2>&1
tail -15 gpurun_out/h_pytest.txt
: > gpurun_out/h_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_FUSED=0" "LSB200_X=0" "LSB200_TMA=0" "LSB200_FUSED_CFG=3" "LSB200_FUSED_CFG=3 LSB200_TMA=0"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/h_matvec.jsonl 2>> gpurun_out/h_matvec.err
  done
done
cut -c1-170 gpurun_out/h_matvec.jsonl
for env in "LSB200_FUSED=0" "LSB200_X=0" "LSB200_TMA=0"; do
  echo "== $env" >> gpurun_out/h_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/h_slab_1024.txt 2>&1
done
cat gpurun_out/h_slab_1024.txt
timeout 600 python bench.py > gpurun_out/h_bench_256.json 2> gpurun_out/h_bench_256.err
tail -3 gpurun_out/h_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/h_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline'], d['roofline_iteration']['frac'], d['e2e'], {k:v for k,v in d['eval'].items() if 'host' in k}, d.get('kernels'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/h_launches_bench_256.csv \
  python bench.py --steps 1 --warmup 1 --no-experiments --no-cpu-baseline --no-fault > gpurun_out/h_ncu_bench.log 2>&1
tail -2 gpurun_out/h_ncu_bench.log | cut -c1-300
