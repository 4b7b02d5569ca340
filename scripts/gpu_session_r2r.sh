#!/bin/bash
# Round 2, GPU session R (TWO B200s): slab parity tests over NCCL (FDI, evaluation after a slab solve, P1 slabs), the
# 2-GPU bench line with its correctness check and a configs[4]-shaped sub-run at 512^3.
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_slabs_gpu.py -m gpu -q -p no:cacheprovider -k "2" > gpurun_out/r_pytest_slabs.txt 2>&1
tail -6 gpurun_out/r_pytest_slabs.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 3 --warmup 3 --c5 --c5-n 512 --eval-points 20000000 > gpurun_out/r_bench_2gpu.json 2> gpurun_out/r_bench_2gpu.err
tail -2 gpurun_out/r_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline']['frac'], d['roofline_iteration']['frac'])
print('check', d.get('check')); print('c5', d.get('c5')); print('kernels', d.get('kernels'))
PY
LSB200_PEER_SCALARS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-check > gpurun_out/r_bench_2gpu_allreduce.json 2> gpurun_out/r_bench_2gpu_allreduce.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench_2gpu_allreduce.json').read().strip().splitlines()[-1])
print('ncclAllReduce scalars:', {k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline_iteration'])
PY
