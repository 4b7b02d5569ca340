#!/bin/bash
# Round 2, GPU session AD (EIGHT B200s): the 8- and 4-GPU bench lines with the one-warp scalar pulls (configs[4] was
# measured in session AA)
set -u
mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2972$n \
  bench.py --gpus $n --steps 3 --warmup 3 --no-c5 > gpurun_out/ad_bench_${n}gpu.json 2> gpurun_out/ad_bench_${n}gpu.err
echo "n=$n rc=$?"
python - $n <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/ad_bench_{sys.argv[1]}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','cg_iterations')}, 'us/it in solve', d['solve_seconds']*1e6/d['cg_iterations'], d['roofline_iteration']['seconds_per_iteration'], d.get('check',{}).get('residual_ok'))
PY
done
