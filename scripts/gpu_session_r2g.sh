#!/bin/bash
set -u
mkdir -p gpurun_out
python -m loopstructural_b200.build > gpurun_out/g_build.txt 2>&1
timeout 700 python scripts/sweep_fused.py 256 > gpurun_out/g_sweep_256.txt 2>&1
tail -1 gpurun_out/g_sweep_256.txt
timeout 900 python scripts/sweep_fused.py 512 > gpurun_out/g_sweep_512.txt 2>&1
tail -1 gpurun_out/g_sweep_512.txt
timeout 900 python scripts/sweep_fused.py 1024 0 128 > gpurun_out/g_sweep_slab_end.txt 2>&1
tail -1 gpurun_out/g_sweep_slab_end.txt
