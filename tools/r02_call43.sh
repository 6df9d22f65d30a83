#!/bin/bash
# last GPU call of the round: full GPU suite + smoke on the final (rebuilt) library
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_last.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
