#!/bin/bash
# session-2 baseline: full GPU suite, Conv3d_1a micro-bench (ones-slot weight gradient), default bench, per-layer step profile
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_e.log; cat gpurun_out/r02_pytest_gpu_e.log
timeout 300 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench3.txt 2>&1; cat gpurun_out/r02_conv1a_bench3.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_e.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
timeout 300 python tools/step_profile.py > gpurun_out/r02_step_profile_e.txt 2>&1; head -40 gpurun_out/r02_step_profile_e.txt
