#!/bin/bash
# small FDI systems: where an iteration of the persistent loop goes
mkdir -p gpurun_out
timeout 200 python scripts/small_systems.py claudius fdi64 > gpurun_out/ap_small.txt 2>&1; cut -c1-700 gpurun_out/ap_small.txt
timeout 100 python scripts/cta_trace.py fdi 56 > gpurun_out/ap_trace56.txt 2>&1; cat gpurun_out/ap_trace56.txt
LSB200_PERSIST=0 timeout 200 python scripts/small_systems.py claudius 2>&1 | cut -c1-400
