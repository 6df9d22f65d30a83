#!/bin/bash
# final-code evidence at N=1: smoke, the driver's default command, reference arm, the optional early-Adam path through the
# trajectory / checkpoint tests, launch list of one step
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/r02_last_bench_n1.json 2> gpurun_out/r02_last_bench_n1.err ) 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_last_bench_reference.json 2> gpurun_out/r02_last_bench_reference.err; echo "reference rc=$?"
OTAL_EARLY_ADAM=1 timeout 300 python -m pytest tests/test_model_b8_gpu.py tests/test_checkpoint_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import json
for f in ("n1", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/r02_last_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 3), d.get("e2e") and round(d["e2e"]["value"], 1), d.get("roofline") and round(d["roofline"]["frac"], 4), d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
        if f == "n1":
            print([(p.get("mode"), p.get("clips_per_s", p.get("error"))) for p in d["other_configs"]["cliplen_batch_sweep"]])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_last.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_last.csv 1 > gpurun_out/r02_launches_last_summary.txt 2>&1; head -12 gpurun_out/r02_launches_last_summary.txt
