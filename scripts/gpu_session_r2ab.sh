#!/bin/bash
# Round 2, GPU session AB (TWO B200s): one-warp pulls of the scalar reductions -- slab parity, 2-GPU bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_slabs_gpu.py -m gpu -q -p no:cacheprovider -k "solve_matches and 2" > gpurun_out/ab_pytest_slabs.txt 2>&1
tail -3 gpurun_out/ab_pytest_slabs.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/ab_bench_2gpu.json 2> gpurun_out/ab_bench_2gpu.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ab_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, 'us/it in solve', d['solve_seconds']*1e6/d['cg_iterations'], d['roofline_iteration']['seconds_per_iteration'], d.get('check',{}).get('residual_ok'), d.get('check',{}).get('vs_single_gpu_solve'))
PY
