#!/bin/bash
# Round 2, GPU session A (one B200): plug-in on the device through the reference's call chain, full-size goldens, the new
# row path and host staging, the bench line with the real reference arm, and per-kernel ncu (time + DRAM bytes).
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash scripts/gpu_session_r2a.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
python -m loopstructural_b200.build > gpurun_out/a_build.txt 2>&1
timeout 1000 python -m pytest tests -q -m gpu --maxfail=15 -p no:cacheprovider > gpurun_out/a_pytest_gpu.txt 2>&1
tail -5 gpurun_out/a_pytest_gpu.txt
timeout 500 python bench.py --steps 5 --warmup 3 --no-experiments > gpurun_out/a_bench_256.json 2> gpurun_out/a_bench_256.err
tail -c 600 gpurun_out/a_bench_256.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
tail -c 400 gpurun_out/a_bench_ref.json
# per-kernel time and DRAM traffic: the CG iteration kernels (first 300 launches after the set-up) ...
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/a_ncu_iteration_256.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-experiments --no-fault > /dev/null 2>&1
# ... and every other kernel of the bench step (set-up, evaluation, fault frames, lattice)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
  -k 'regex:k_eval|k_fault|k_frame|k_field|k_fdi_point|k_fdi_pair|k_rows_rhs|k_rows_diag|k_rows_ingest|k_rows_sort|k_rows_permute|k_fdi_diag|k_scaling|k_cg_residual|k_p1' \
  --log-file gpurun_out/a_ncu_other_256.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-experiments > /dev/null 2>&1
# the row kernel in full
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rows_normal -s 4 -c 2 -o gpurun_out/a_rows_full \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-experiments --no-fault > /dev/null 2>&1
ls -la gpurun_out | tail -15
