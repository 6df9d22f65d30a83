#!/bin/bash
# 1 GPU: memcheck of the half-precision BoundaryMaxPooling tests, then the full GPU suite
set -u
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bmp_gpu.py -q -x -k half 2>&1 | tail -12 | tee gpurun_out/r02c_memcheck_bmp_half.log
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_last.log
