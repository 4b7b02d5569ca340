#!/bin/bash
# Round 2, GPU session F (one B200): k_fdi_fused with interior march + z-face gather + shared strip routine; TMA staging.
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/f_build.txt 2>&1
timeout 300 python -m pytest tests/test_fdi_gpu.py -q -p no:cacheprovider > gpurun_out/f_pytest.txt 2>&1
tail -2 gpurun_out/f_pytest.txt
LSB200_TMA=0 timeout 300 python -m pytest tests/test_fdi_gpu.py -q -p no:cacheprovider -x > gpurun_out/f_pytest_notma.txt 2>&1
tail -2 gpurun_out/f_pytest_notma.txt
: > gpurun_out/f_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_FUSED=0" "" "LSB200_TMA=0" "LSB200_FUSED_CFG=3" "LSB200_FUSED_CFG=3 LSB200_TMA=0" "LSB200_FUSED_SCZ=16" "LSB200_FUSED_SCZ=64"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/f_matvec.jsonl 2>> gpurun_out/f_matvec.err
  done
done
cut -c1-150 gpurun_out/f_matvec.jsonl
for env in "LSB200_FUSED=0" "" "LSB200_TMA=0"; do
  echo "== $env" >> gpurun_out/f_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/f_slab_1024.txt 2>&1
done
cat gpurun_out/f_slab_1024.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o gpurun_out/f_fused_256 \
  python scripts/matvec_experiment.py 256 > /dev/null 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-experiments --no-cpu-baseline --no-fault > gpurun_out/f_bench_256.json 2> gpurun_out/f_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench_256.json'))
print({k:d[k] for k in ('ms_per_step','cg_iterations')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['eval'].get('host_arrays'), d['eval'].get('gradient_points_per_second_host_arrays'))
PY
