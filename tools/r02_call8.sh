#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_backbone_kernels_gpu.py -m gpu -q -x -k pool 2>&1 | tail -2
timeout 200 python tools/pool_bench.py > gpurun_out/r02_pool_bench.txt 2>&1; cat gpurun_out/r02_pool_bench.txt
