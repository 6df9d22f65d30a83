#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backbone_kernels_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head
timeout 200 python tools/pool_bench.py > gpurun_out/r02_pool_bench_c.txt 2>&1; grep fused gpurun_out/r02_pool_bench_c.txt | sed 's/.*bwd fused/rows bwd fused/'
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_pool_bench_step.json 2> gpurun_out/r02_pool_bench_step.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_pool_bench_step.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'])
PY
