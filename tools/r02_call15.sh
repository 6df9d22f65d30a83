#!/bin/bash
# explicit head schedule: its own tests, the GPU suite, A/B bench (OTAL_HEAD_SCHEDULE=0 vs default), launch list
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head_schedule_gpu.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-400 | head -30
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head -30
for v in 0 1; do
OTAL_HEAD_SCHEDULE=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_hs_bench_$v.json 2> gpurun_out/r02_hs_bench_$v.err; echo "bench sched=$v rc=$?"; tail -2 gpurun_out/r02_hs_bench_$v.err
done
python - <<PY
import json
for v in (0, 1):
    try:
        d=json.loads(open(f'gpurun_out/r02_hs_bench_{v}.json').read().strip().splitlines()[-1])
        print(v, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'])
    except Exception as e: print(v, 'failed', e)
PY
