#!/bin/bash
# persistent loop: barrier on one monotonic counter
mkdir -p gpurun_out
timeout 120 python scripts/small_systems.py claudius fdi64 > gpurun_out/as_small.txt 2>&1; cut -c1-420 gpurun_out/as_small.txt
LSB200_P1_MARCH=3 timeout 120 python scripts/p1_experiment.py 117 2>&1 | cut -c1-200
timeout 400 python -m pytest tests/test_zz_claudius_gpu.py tests/test_p1_gpu.py tests/test_fdi_gpu.py -x -q -k "not full_size" 2>&1 | tail -3
