#!/bin/bash
# N GPUs: how many CTAs should the overlapped gradient all-reduce take from the backward?  (NCCL_MAX_CTAS A/B on the default step)
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for c in default 8 4 2 default 8 4 2; do
  if [ "$c" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  timeout 200 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NCCL_MAX_CTAS=$c', d['n_gpus'], round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms', d['dp_params_in_sync'])
"
done 2>&1 | tee gpurun_out/r02_nccl_ctas_n$N.txt
