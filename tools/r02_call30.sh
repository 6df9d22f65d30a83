#!/bin/bash
# 1 GPU: re-capture of the step graph on one Trainer (the sweep's pattern) — with the persistent capture stream and with a fresh one
set -u
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 150 python tools/probe/recapture_dp.py 2>&1 | grep -v "SyntaxWarning\|logit: softmax\|^$" | grep " B \|Error\|in sync" | head -12; }
{
run persist_stream PROBE_MODE=destroy
run new_stream OTAL_CAP_STREAM=new PROBE_MODE=destroy
} 2>&1 | tee gpurun_out/r02_recapture_probe_n1.txt
# the sweep leg of the default bench line, as the driver runs it
( time timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_full_bench_n1.json 2> gpurun_out/r02_full_bench_n1.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_full_bench_n1.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'])
o=d['other_configs']
print(o['anet']); print(o['inference'])
print([(p.get('mode'), p.get('clips_per_s', p.get('error'))) for p in o['cliplen_batch_sweep']])
PY
