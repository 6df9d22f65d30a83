#!/bin/bash
# launch list of the current step (ncu, serialised) + summary
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r02_launches_b_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_b.csv 1 > gpurun_out/r02_launches_b_summary.txt 2>&1; head -70 gpurun_out/r02_launches_b_summary.txt
