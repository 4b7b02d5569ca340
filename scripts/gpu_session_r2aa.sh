#!/bin/bash
# Round 2, GPU session AA (EIGHT B200s): slab parity tests on 8 ranks, the 8-GPU bench line (weak scaling of configs[1],
# correctness check) with the configs[4] sub-object (1024^3 solve + 1e8-point replicated evaluation)
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_slabs_gpu.py -m gpu -q -p no:cacheprovider -k "8" > gpurun_out/aa_pytest_slabs.txt 2>&1
tail -4 gpurun_out/aa_pytest_slabs.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 \
  bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/aa_bench_8gpu.json 2> gpurun_out/aa_bench_8gpu.err
echo "rc=$?"; tail -2 gpurun_out/aa_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/aa_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline']['frac'], d['roofline_iteration'])
print('check', d.get('check')); print('c5', d.get('c5')); print('kernels', d.get('kernels'))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 \
  bench.py --gpus 4 --steps 3 --warmup 3 --no-check > gpurun_out/aa_bench_4gpu.json 2> gpurun_out/aa_bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/aa_bench_4gpu.json').read().strip().splitlines()[-1])
print('N=4', {k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline_iteration']['seconds_per_iteration'])
PY
