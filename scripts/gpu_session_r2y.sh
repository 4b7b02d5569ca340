#!/bin/bash
# Round 2, GPU session Y (TWO B200s): the 2-GPU bench line (peer-mapped scalar reductions) with its correctness check and a
# configs[4]-shaped sub-run at 512^3
set -u
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 2 --warmup 3 --c5 --c5-n 512 --eval-points 20000000 > gpurun_out/y_bench_2gpu.json 2> gpurun_out/y_bench_2gpu.err
echo "rc=$?"; tail -2 gpurun_out/y_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline']['frac'], d['roofline_iteration'])
print('check', d.get('check')); print('c5', d.get('c5')); print('kernels', d.get('kernels'))
PY
