#!/bin/bash
# Round 2, GPU session C (one B200): k_fdi_fused after the staging fix, chunk sweeps; failed tests of session B.
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/c_build.txt 2>&1
timeout 600 python -m pytest tests/test_zz_fullsize_gpu.py tests/test_zz_plugin_device_gpu.py tests/test_fdi_gpu.py -q -p no:cacheprovider > gpurun_out/c_pytest.txt 2>&1
tail -4 gpurun_out/c_pytest.txt
: > gpurun_out/c_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_FUSED=0" "" "LSB200_FUSED_CFG=3" "LSB200_FUSED_SCZ=16" "LSB200_FUSED_SCZ=32" "LSB200_FUSED_SCZ=64"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/c_matvec.jsonl 2>> gpurun_out/c_matvec.err
  done
done
cat gpurun_out/c_matvec.jsonl | cut -c1-200
for env in "LSB200_FUSED=0" ""; do
  echo "== $env" >> gpurun_out/c_slab_1024.txt
  timeout 300 env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/c_slab_1024.txt 2>&1
done
cat gpurun_out/c_slab_1024.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fdi_fused -s 3 -c 1 -o gpurun_out/c_fused_256 \
  python scripts/matvec_experiment.py 256 > /dev/null 2>&1
ls -la gpurun_out | tail -6
