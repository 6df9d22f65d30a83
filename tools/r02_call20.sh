#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv1a_u8_gpu.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | cut -c1-300 | head -20
timeout 200 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench4.txt 2>&1; cat gpurun_out/r02_conv1a_bench4.txt | tail -14
