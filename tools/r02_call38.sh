#!/bin/bash
# 1 GPU: K-tail trimming in conv_igemm (zero K steps of a tap's last chunk not issued): GPU suite, then A/B/A/B
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_last.log
for e in 1 0 1 0; do
  if [ $e = 1 ]; then export OTAL_NO_KTAIL=1; else unset OTAL_NO_KTAIL; fi
  timeout 200 python bench.py --steps 40 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('OTAL_NO_KTAIL=$e', round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms; conv_igemm', round(r['ms_per_step'], 3), 'ms frac', round(r['frac'], 4))
"
done 2>&1 | tee gpurun_out/r02_ktail_ab.txt
