#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head_schedule_gpu.py tests/test_model_gpu.py tests/test_model_b8_gpu.py tests/test_model_anet_gpu.py tests/test_head_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-300 | head
cat > /tmp/ab.py <<'PY'
import os, sys, json, subprocess
for tag, env in (("two", {}), ("one", {"OTAL_HEAD_ONE_STREAM": "1"}), ("two", {}), ("one", {"OTAL_HEAD_ONE_STREAM": "1"})):
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3", "--no-cpu-baseline", "--no-other-configs", "--no-e2e"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1]); print(tag, round(d["value"], 1), round(d["ms_per_step"], 3))
    except Exception as ex:
        print(tag, "failed", out.stderr[-500:])
PY
python /tmp/ab.py
