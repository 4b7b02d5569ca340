#!/bin/bash
# Round 2, GPU session M (one B200): full device suite after the prune, matvec plans, P1 tiles, default bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/m_pytest.txt 2>&1
tail -4 gpurun_out/m_pytest.txt
: > gpurun_out/m_matvec.jsonl
for n in 256 512; do
  for env in "LSB200_X=0" "LSB200_TMA=0" "LSB200_FUSED_CFG=3"; do
    timeout 200 env $env python scripts/matvec_experiment.py $n >> gpurun_out/m_matvec.jsonl 2>> gpurun_out/m_matvec.err
  done
done
cut -c1-150 gpurun_out/m_matvec.jsonl
timeout 300 python scripts/prof_slab.py 1024 128 > gpurun_out/m_slab_1024.txt 2>&1; cat gpurun_out/m_slab_1024.txt
: > gpurun_out/m_p1.jsonl
for env in "LSB200_P1_TILES=0" "LSB200_P1_TILE_CFG=0" "LSB200_P1_TILE_CFG=1" "LSB200_P1_TILE_CFG=2"; do
  timeout 200 env $env python scripts/p1_experiment.py 117 >> gpurun_out/m_p1.jsonl 2>> gpurun_out/m_p1.err
done
cut -c1-130 gpurun_out/m_p1.jsonl
timeout 200 python scripts/cta_trace.py p1 117 > gpurun_out/m_trace_p1_117.txt 2>&1; tail -8 gpurun_out/m_trace_p1_117.txt
timeout 600 python bench.py --no-experiments > gpurun_out/m_bench_256.json 2> gpurun_out/m_bench_256.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/m_bench_256.json'))
print({k:d.get(k) for k in ('ms_per_step','cg_iterations','value')}, d['roofline'], d['roofline_iteration']['frac'], d['e2e'], {k:v for k,v in d['eval'].items() if 'host' in k}, d.get('kernels'))
PY
