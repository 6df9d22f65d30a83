#!/bin/bash
# last GPU call of the round: the full GPU suite and smoke on the final build, one short default-step bench
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_last.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --steps 40 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'], 1), 'clips/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 1), 'roofline', round(d['roofline']['frac'], 4), d['clocks'])
" | tee gpurun_out/r02_last_short_bench.txt
