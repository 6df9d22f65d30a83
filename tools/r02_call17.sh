#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head -30
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_hs_bench_2.json 2> gpurun_out/r02_hs_bench_2.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_hs_bench_2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_hs_bench_2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'])
PY
OTAL_NO_WGRAD_OVERLAP=1 timeout 300 python tools/step_profile.py --u8 --top 40 > gpurun_out/r02_step_profile_h.txt 2>&1; head -36 gpurun_out/r02_step_profile_h.txt
