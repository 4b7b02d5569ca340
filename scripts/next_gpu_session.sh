#!/bin/bash
# First GPU call of the next session (one B200): everything that was written after round 1's GPU budget was spent, in the
# order that matters if the call is cut short.  Results land in gpurun_out/ (copy what is worth keeping to profiles/).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/next_gpu_session.sh'
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/n_build.txt 2>&1
# 1. the launch plans that are built but unmeasured: fused ring / staggered ring / shell first, 256^3 and 512^3
for n in 256 512; do
  for env in "" "LSB200_RING=1" "LSB200_RING=2" "LSB200_SHELL_FIRST=1"; do
    env $env python scripts/matvec_experiment.py $n >> gpurun_out/n_launch_plans.jsonl 2>> gpurun_out/n_launch_plans.err
  done
done
# one end-rank and one middle-rank slab of 1024^3 (the per-GPU share of BASELINE configs[4]) under the same plans
for env in "" "LSB200_RING=1" "LSB200_RING=2"; do
  echo "== $env" >> gpurun_out/n_slab_1024.txt
  env $env python scripts/prof_slab.py 1024 128 >> gpurun_out/n_slab_1024.txt 2>&1
done
# 2. the device tests that have never run: opt-in plans + slab handles on one device, P1 slabs, fault frames / lattice / rows
for env in "" "LSB200_P1_FAST=1"; do env $env python scripts/p1_experiment.py 117 >> gpurun_out/n_launch_plans.jsonl 2>> gpurun_out/n_launch_plans.err; done
python -m pytest tests/test_zz_d_ring_gpu.py tests/test_zz_e_p1_slab_gpu.py tests/test_zz_fault_gpu.py -q > gpurun_out/n_pytest_new.txt 2>&1
# 3. the whole device suite and the bench line (with the "fault" and "launch_plan_experiments" keys)
python -m pytest tests -q -m gpu > gpurun_out/n_pytest_gpu.txt 2>&1
python bench.py > gpurun_out/n_bench_256.json 2> gpurun_out/n_bench_256.err
# 4. ncu: launch list of one bench step under the fused ring, full capture of the fused box kernel and of the fault kernel
LSB200_RING=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_ring.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fault --no-experiments > /dev/null 2>&1
LSB200_RING=1 ncu --set full --clock-control none --import-source on -k regex:k_fdi_box -s 6 -c 2 -o gpurun_out/n_box_ring \
  python scripts/matvec_experiment.py 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fault_displace -c 1 -o gpurun_out/n_fault \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-experiments > /dev/null 2>&1
ls -la gpurun_out | tail -20
