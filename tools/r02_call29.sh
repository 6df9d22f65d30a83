#!/bin/bash
# 2 GPUs: re-capture of the data-parallel step graph on one Trainer — the fallback (exchange behind the replay) and three probes
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
run() { echo "== $1"; shift; env "$@" timeout 150 $TR tools/probe/recapture_dp.py 2>&1 | grep -v "^W1017\|OMP_NUM\|^\*\*\*\|SyntaxWarning\|logit: softmax\|^$" | grep " B \|Error\|in sync" | head -12; }
{
run fallback PROBE_MODE=destroy
run recapture_persist OTAL_DP_RECAPTURE=1 PROBE_MODE=destroy
run recapture_persist_nomix OTAL_DP_RECAPTURE=1 NCCL_GRAPH_MIXING_SUPPORT=0 PROBE_MODE=destroy
run recapture_new_nomix OTAL_DP_RECAPTURE=1 OTAL_CAP_STREAM=new NCCL_GRAPH_MIXING_SUPPORT=0 PROBE_MODE=destroy
} 2>&1 | tee gpurun_out/r02_recapture_probe2.txt
