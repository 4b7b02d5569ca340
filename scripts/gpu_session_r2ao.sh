#!/bin/bash
# config 3: which kernels make up the 60 us iteration
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/ao_launches_c3.csv \
  python scripts/small_systems.py p1 > gpurun_out/ao_ncu_c3.log 2>&1
python scripts/ncu_summary.py gpurun_out/ao_launches_c3.csv > gpurun_out/ao_c3_summary.csv 2>&1; head -20 gpurun_out/ao_c3_summary.csv
