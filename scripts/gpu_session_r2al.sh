#!/bin/bash
# fold rows generated on the device: kernel parity against the reference's families, the refolded model through the plug-in
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_p1_gpu.py tests/test_zz_plugin_device_gpu.py tests/test_zz_iso_gpu.py -x -q > gpurun_out/r2al_tests.txt 2>&1
tail -6 gpurun_out/r2al_tests.txt
LSB200_P1_MARCH=3 timeout 120 python scripts/p1_experiment.py 117 > gpurun_out/r2al_p1.txt 2>&1; cut -c1-400 gpurun_out/r2al_p1.txt
timeout 200 python scripts/small_systems.py p1 > gpurun_out/r2al_small_p1.txt 2>&1; tail -5 gpurun_out/r2al_small_p1.txt
