#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/conv_timeline.py > gpurun_out/r02_conv_timeline2.txt 2>&1; echo "timeline rc=$?"; grep "===\|mean tile" gpurun_out/r02_conv_timeline2.txt
python -m pytest tests/test_conv_gpu.py tests/test_backbone_kernels_gpu.py tests/test_conv1a_u8_gpu.py tests/test_head_gpu.py -m gpu -q -x 2>&1 | tail -5
