#!/bin/bash
# 2 GPUs: early optimizer launches A/B/A/B on the data-parallel step
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
for e in 0 1 0 1; do
  OTAL_EARLY_ADAM=$e timeout 200 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('OTAL_EARLY_ADAM=$e', d['n_gpus'], round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms', d['dp_params_in_sync'])
"
done 2>&1 | tee gpurun_out/r02_early_adam_n2.txt
