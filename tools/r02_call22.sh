#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1 0 1; do
OTAL_CONV_KSPLIT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02_ks_bench_$v.json 2> gpurun_out/r02_ks_bench_$v.err; echo "bench ksplit=$v rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ks_bench_$v.json').read().strip().splitlines()[-1])
print($v, round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'])
PY
done
OTAL_CONV_KSPLIT=1 timeout 300 python -m pytest tests/test_head_schedule_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head
