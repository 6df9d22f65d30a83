#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_d.log
for cfg in "" "--mode infer" "--config anet" "--frames 512 --batch 4" "--mode infer --config anet --batch 4"; do
  tag=$(echo "$cfg" | tr -d ' -'); tag=${tag:-default}
  timeout 600 python bench.py --steps 20 --warmup 3 $cfg > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err
  echo "bench [$cfg] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/r02_bench_${tag}.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'] if d.get('e2e') else None, (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'))" 2>&1 | tail -1)"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; head -30 gpurun_out/r02_launches_summary.txt
