#!/bin/bash
# 2 GPUs: re-capture of the data-parallel step graph on one Trainer (in-graph exchange, the default, and the developer switch that
# keeps it behind the replay), global normalisers under NCCL, then the driver's launch line with every leg
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
run() { echo "== $1"; shift; env "$@" timeout 150 $TR tools/probe/recapture_dp.py 2>&1 | grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax\|^$" | grep " B \|Error\|in sync" | sort | head -12; }
{
run recapture_in_graph_default PROBE_MODE=destroy
run exchange_behind_replay OTAL_DP_RECAPTURE=0 PROBE_MODE=destroy
} 2>&1 | tee gpurun_out/r02_recapture_probe_n2.txt
echo "== global normalisers under NCCL"
timeout 200 python tools/probe/global_norm_nccl.py 2>&1 | grep -v "SyntaxWarning\|logit: softmax" | tail -3
timeout 200 $TR tools/probe/global_norm_nccl.py 2>&1 | grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax\|^$" | tail -30 | tee gpurun_out/r02_global_norm_nccl.txt
( time timeout 500 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_full_bench_n2.json 2> gpurun_out/r02_full_bench_n2.err ) 2>&1 | tail -3
grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax" gpurun_out/r02_full_bench_n2.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_full_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['dp_params_in_sync'], d['gpu_launches'])
o=d['other_configs']
print(o['anet']); print(o['inference'])
print([(p.get('mode'), p.get('clips_per_s', p.get('error'))) for p in o['cliplen_batch_sweep']])
PY
