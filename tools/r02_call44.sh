#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_model_b8_gpu.py tests/test_checkpoint_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | tail -2
OTAL_NVTX=1 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok (NVTX on)')" 2>&1 | tail -1
