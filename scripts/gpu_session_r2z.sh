#!/bin/bash
# Round 2, GPU session Z (TWO B200s): halo planes of p over peer memory -- slab parity tests, 2-GPU bench against the
# ncclSend/Recv exchange
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_slabs_gpu.py -m gpu -q -p no:cacheprovider -k "2" > gpurun_out/z_pytest_slabs.txt 2>&1
tail -4 gpurun_out/z_pytest_slabs.txt
for tag in peer nccl; do
  if [ $tag = nccl ]; then export LSB200_PEER_HALO=0; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/z_bench_2gpu_$tag.json 2> gpurun_out/z_bench_2gpu_$tag.err
  echo "$tag rc=$?"
  python - $tag <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/z_bench_2gpu_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, d['roofline_iteration']['seconds_per_iteration'], d.get('check',{}).get('residual_ok'), d.get('check',{}).get('vs_single_gpu_solve'))
PY
done
