#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 200 python tools/probe/global_norm_nccl.py 2>&1 | grep -v "SyntaxWarning\|logit: softmax" | tail -3
timeout 200 $TR tools/probe/global_norm_nccl.py > gpurun_out/r02_global_norm_nccl.txt 2>&1
grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax\|^$" gpurun_out/r02_global_norm_nccl.txt | grep -B30 "ChildFailedError" | head -60
grep "eager\|graph\|terms\|IoU\|ok" gpurun_out/r02_global_norm_nccl.txt | cut -c1-600
