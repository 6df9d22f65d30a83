#!/bin/bash
# round-2 evidence at N=1: default bench line (incl. cpu_baseline + other_configs), reference arm, infer / anet / ssl modes, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference rc=$?"
timeout 300 python bench.py --mode infer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_infer.json 2>/dev/null; echo "infer rc=$?"
timeout 300 python bench.py --ssl --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02_final_bench_ssl.json 2>/dev/null; echo "ssl rc=$?"
timeout 300 python bench.py --config anet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_anet.json 2>/dev/null; echo "anet rc=$?"
python - <<'PY'
import json
for f in ("n1", "reference", "infer", "ssl", "anet"):
    try:
        d = json.loads(open(f"gpurun_out/r02_final_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 3), d.get("e2e") and round(d["e2e"]["value"], 1), d.get("roofline") and round(d["roofline"]["frac"], 3), d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_final.csv 1 > gpurun_out/r02_launches_final_summary.txt 2>&1; head -50 gpurun_out/r02_launches_final_summary.txt
