#!/bin/bash
# Round 2, GPU session AC (one B200): P1 orthogonality rows and the plug-in on the device; k_p1_tiles with row-wise staging
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_fault_gpu.py tests/test_zz_plugin_device_gpu.py tests/test_p1_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/ac_pytest.txt 2>&1
tail -5 gpurun_out/ac_pytest.txt
: > gpurun_out/ac_p1.jsonl
for env in "LSB200_P1_TILE_CFG=0" "LSB200_P1_TILE_CFG=1" "LSB200_P1_TILE_CFG=2"; do
  timeout 200 env $env python scripts/p1_experiment.py 117 >> gpurun_out/ac_p1.jsonl 2>> gpurun_out/ac_p1.err
done
cut -c1-130 gpurun_out/ac_p1.jsonl
timeout 300 python scripts/small_systems.py p1 claudius | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['system'], 'its', d['iterations'], 'solve_ms %.2f' % d['solve_ms'], 'us/it %.1f' % d['us_per_iteration_solve'], 'matvec %.1f' % d['matvec_us'])"
