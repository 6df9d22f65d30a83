#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv1a|conv_igemm|conv_wgrad|maxpool|gn_relu|msl_' -c 16 \
  -o gpurun_out/r02_kernels -f python tools/ncu_targets_r02.py > gpurun_out/r02_ncu_targets.log 2>&1; echo "ncu rc=$?"
tail -22 gpurun_out/r02_ncu_targets.log
ls -la gpurun_out/r02_kernels.ncu-rep
