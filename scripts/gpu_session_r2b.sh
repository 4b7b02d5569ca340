#!/bin/bash
# Round 2, GPU session B (one B200): the single-launch FDI matvec (k_fdi_fused) against the three-launch plan of round 1.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_session_r2b.sh'
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/b_build.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --maxfail=15 -p no:cacheprovider > gpurun_out/b_pytest_gpu.txt 2>&1
tail -4 gpurun_out/b_pytest_gpu.txt
: > gpurun_out/b_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_FUSED=0" "" "LSB200_FUSED_CFG=3" "LSB200_FUSED_CFG=4" "LSB200_FUSED_SCZ=32" "LSB200_FUSED_SCZ=128"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/b_matvec.jsonl 2>> gpurun_out/b_matvec.err
  done
done
for cz in 32 37 52 64 86; do
  timeout 200 env LSB200_FUSED_CZ=$cz python scripts/matvec_experiment.py 256 >> gpurun_out/b_matvec.jsonl 2>> gpurun_out/b_matvec.err
done
cat gpurun_out/b_matvec.jsonl | cut -c1-260
for env in "LSB200_FUSED=0" ""; do
  echo "== $env" >> gpurun_out/b_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/b_slab_1024.txt 2>&1
done
cat gpurun_out/b_slab_1024.txt
timeout 500 python bench.py --steps 5 --warmup 3 --no-experiments --no-cpu-baseline > gpurun_out/b_bench_256.json 2> gpurun_out/b_bench_256.err
tail -c 1500 gpurun_out/b_bench_256.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o gpurun_out/b_fused_256 \
  python scripts/matvec_experiment.py 256 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o gpurun_out/b_fused_512 \
  python scripts/matvec_experiment.py 512 > /dev/null 2>&1
ls -la gpurun_out | tail -12
