#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/x_plan_check.jsonl
for grid in "32 40 27" "55 68 47"; do
  for env in "LSB200_X=0" "LSB200_AUTOTUNE=0" "LSB200_TMA=0" "LSB200_FUSED_CZ=2 LSB200_FUSED_SCZ=3" "LSB200_FUSED_CZ=3 LSB200_FUSED_SCZ=4" "LSB200_FUSED_CZ=1 LSB200_FUSED_SCZ=2" "LSB200_FUSED_CZ=4 LSB200_FUSED_SCZ=6 LSB200_TMA=0" "LSB200_FUSED_CZ=2 LSB200_FUSED_SCZ=3 LSB200_TMA=0"; do
    timeout 120 env $env python scripts/plan_check.py $grid >> gpurun_out/x_plan_check.jsonl 2>> gpurun_out/x_plan_check.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/x_plan_check.jsonl'):
    d=json.loads(l); print(d['grid'], d['env'], '%.2e' % d['max_rel_diff'], '%.2e' % d['repeat_diff'], d['plan'][:140])
PY
tail -3 gpurun_out/x_plan_check.err
timeout 600 python -m pytest tests/test_zz_claudius_gpu.py tests/test_zz_plugin_device_gpu.py -m gpu -q -p no:cacheprovider -k "claudius or readme" > gpurun_out/x_pytest.txt 2>&1; tail -4 gpurun_out/x_pytest.txt
LSB200_PERSIST=0 timeout 600 python -m pytest tests/test_zz_claudius_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/x_pytest_nopersist.txt 2>&1; tail -3 gpurun_out/x_pytest_nopersist.txt
LSB200_AUTOTUNE=0 timeout 600 python -m pytest tests/test_zz_claudius_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/x_pytest_notune.txt 2>&1; tail -3 gpurun_out/x_pytest_notune.txt
