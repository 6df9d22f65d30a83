#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv1a_u8_gpu.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | cut -c1-300 | head -20
timeout 200 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench4.txt 2>&1; cat gpurun_out/r02_conv1a_bench4.txt | tail -14
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head -30
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_g.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], round(d['roofline']['frac'],3))
PY
