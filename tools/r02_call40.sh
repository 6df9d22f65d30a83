#!/bin/bash
# 1 GPU: third pipeline stage + one staging buffer instead of 2 + 2 where it fits (BN <= 112 in bf16x3): A/B/A/B, conv tests under the switch
set -u
mkdir -p gpurun_out
OTAL_CONV_PREFER_STAGES=1 python -m pytest tests/test_conv_gpu.py tests/test_backbone_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -2
for e in 0 1 0 1; do
  if [ $e = 1 ]; then export OTAL_CONV_PREFER_STAGES=1; else unset OTAL_CONV_PREFER_STAGES; fi
  timeout 200 python bench.py --steps 40 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('OTAL_CONV_PREFER_STAGES=$e', round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms; conv_igemm', round(r['ms_per_step'], 3), 'ms frac', round(r['frac'], 4))
"
done 2>&1 | tee gpurun_out/r02_prefer_stages_ab.txt
