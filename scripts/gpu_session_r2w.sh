#!/bin/bash
# Round 2, GPU session W (one B200): full device suite with the even row pitch
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/w_pytest.txt 2>&1
tail -12 gpurun_out/w_pytest.txt
