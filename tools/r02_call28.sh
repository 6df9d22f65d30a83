#!/bin/bash
# the driver's launch line at N GPUs with every leg (default bench incl. other_configs), tight timeout
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
( time timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_full_bench_n$N.json 2> gpurun_out/r02_full_bench_n$N.err ) 2>&1 | tail -3; echo "bench N=$N rc=$?"
grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax" gpurun_out/r02_full_bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_full_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['dp_params_in_sync'], d['gpu_launches'])
o=d['other_configs']
print(o['anet']); print(o['inference'])
print(len(o['cliplen_batch_sweep']), [p.get('clips_per_s', p.get('error')) for p in o['cliplen_batch_sweep']])
PY
