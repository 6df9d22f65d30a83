#!/bin/bash
# in-graph bucketed all-reduce + device-step Adam: GPU suite, then bench at N GPUs
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | cut -c1-300 | head -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_dp_bench_n1.json 2> gpurun_out/r02_dp_bench_n1.err; echo "bench N=1 rc=$?"; tail -3 gpurun_out/r02_dp_bench_n1.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_dp_bench_n$N.json 2> gpurun_out/r02_dp_bench_n$N.err; echo "bench N=$N rc=$?"
tail -5 gpurun_out/r02_dp_bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open(f'gpurun_out/r02_dp_bench_n{n}.json').read().strip().splitlines()[-1])
        print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['dp_params_in_sync'], d['gpu_launches'])
    except Exception as e: print(n, 'failed', e)
PY
