#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 tools/probe/bin/tma_probe > gpurun_out/r02_tma_probe.txt 2>&1; echo "probe rc=$?"
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_b.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2>gpurun_out/r02_bench_b.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_b.json
