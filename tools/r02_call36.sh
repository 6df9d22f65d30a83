#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_bmp_gpu.py -q -x -k half > /tmp/mc.log 2>&1
grep -n "=========" /tmp/mc.log | grep -v "Host Frame" | head -40 | tee gpurun_out/r02c_memcheck_bmp_half.log
grep -n "Host Frame" /tmp/mc.log | head -12
