#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv1a_u8_gpu.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | cut -c1-300 | head -20
timeout 200 python tools/conv1a_bench.py 2>&1 | grep -E "parity|wgrad"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv1a_wgrad_halo' -c 1 -o gpurun_out/r02_conv1a_wgrad_halo -f python tools/ncu_targets_r02.py > gpurun_out/r02_ncu_wgrad.log 2>&1; echo "ncu rc=$?"
