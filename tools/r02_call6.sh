#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_c.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2>gpurun_out/r02_bench_c.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_c.json
OTAL_NO_WGRAD_OVERLAP=1 timeout 400 python tools/step_profile.py --top 600 > gpurun_out/r02_step_profile_c.txt 2>&1; echo "step_profile rc=$?"; head -12 gpurun_out/r02_step_profile_c.txt
