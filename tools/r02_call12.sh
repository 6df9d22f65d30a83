#!/bin/bash
# 8-GPU evidence: the driver's own launch line at N GPUs (default bench incl. the other_configs legs: ActivityNet config, clip length x
# batch sweep, inference protocol) -> gpurun_out/r02_mg_bench_n$N.json
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_mg_bench_n$N.json 2> gpurun_out/r02_mg_bench_n$N.err ) 2>&1 | tail -3; echo "bench N=$N rc=$?"
tail -2 gpurun_out/r02_mg_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_mg_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['dp_params_in_sync'])
o=d['other_configs']
print(o['anet']); print(o['inference'])
for p in o['cliplen_batch_sweep']: print(p)
PY
