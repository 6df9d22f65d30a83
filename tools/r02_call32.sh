#!/bin/bash
# 2 GPUs: global normalisers under NCCL (whole batch on one process first), then the NCCL_MAX_CTAS A/B
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
echo "== global normalisers under NCCL"
timeout 200 python tools/probe/global_norm_nccl.py 2>&1 | grep -v "SyntaxWarning\|logit: softmax" | tail -3
timeout 200 $TR tools/probe/global_norm_nccl.py 2>&1 | grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax\|^$" | tail -40 | tee gpurun_out/r02_global_norm_nccl.txt
bash tools/r02_call31.sh 2
