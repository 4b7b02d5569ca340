#!/bin/bash
# full device suite with the fold / directional rows and the marching P1 operator, then the default bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/an_pytest.txt 2>&1
tail -4 gpurun_out/an_pytest.txt
timeout 900 python bench.py > gpurun_out/an_bench_256.json 2> gpurun_out/an_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/an_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline']['frac'], d['roofline_iteration']['frac'], d['e2e'], d.get('other_configs'), d['config'].get('launch_plan'))
PY
