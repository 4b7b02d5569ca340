#!/bin/bash
# Round 2, GPU session D (one B200): k_fdi_fused with the rare paths out of line (162 registers, no spills).
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/e_build.txt 2>&1
timeout 600 python -m pytest tests/test_fdi_gpu.py -q -p no:cacheprovider > gpurun_out/e_pytest.txt 2>&1
tail -2 gpurun_out/e_pytest.txt
: > gpurun_out/e_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_FUSED=0" "" "LSB200_FUSED_CFG=3" "LSB200_FUSED_SCZ=8" "LSB200_FUSED_SCZ=16" "LSB200_FUSED_SCZ=32" "LSB200_FUSED_CZ=37"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/e_matvec.jsonl 2>> gpurun_out/e_matvec.err
  done
done
cat gpurun_out/e_matvec.jsonl | cut -c1-130
for env in "LSB200_FUSED=0" ""; do
  echo "== $env" >> gpurun_out/e_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/e_slab_1024.txt 2>&1
done
cat gpurun_out/e_slab_1024.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o gpurun_out/e_fusee_256 \
  python scripts/matvec_experiment.py 256 > /dev/null 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-experiments --no-cpu-baseline --no-fault > gpurun_out/e_bench_256.json 2> gpurun_out/e_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/e_bench_256.json'))
print({k:d[k] for k in ('ms_per_step','cg_iterations')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['eval'].get('host_arrays'), d['eval'].get('gradient_points_per_secone_host_arrays'))
PY
