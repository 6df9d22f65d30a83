#!/bin/bash
# fused loss flavours (ActivityNet, closed-set focal) + the default bench with its other_configs legs
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msl_gpu.py -m gpu -q -x 2>&1 | grep -v "^E    " | tail -25
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_f.log; cat gpurun_out/r02_pytest_gpu_f.log
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err ) 2>&1 | tail -3
tail -3 gpurun_out/r02_bench_f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f.json').read().strip().splitlines()[-1])
print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])
o=d['other_configs']
print(o['anet']); print(o['inference'])
for p in o['cliplen_batch_sweep']: print(p)
PY
