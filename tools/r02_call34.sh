#!/bin/bash
# 1 GPU: full GPU suite on the final code, then early-Adam A/B/A/B on the default step
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_last.log
for e in 0 1 0 1; do
  OTAL_EARLY_ADAM=$e timeout 200 python bench.py --steps 40 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('OTAL_EARLY_ADAM=$e', round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms', d['gpu_launches'])
"
done 2>&1 | tee gpurun_out/r02_early_adam_n1.txt
