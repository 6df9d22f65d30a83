#!/bin/bash
# Multi-GPU call of round 2 (charged N x box time — keep it short):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_multi_gpu_call.sh 8'
# 1. bench.py at N GPUs (the driver's own launch line), 2. BASELINE configs[3] (ActivityNet, 768-frame clips) and a slice of
# configs[4] (clip length x batch) data parallel through tools/sweep.py, 3. the synthetic-dataset training run on 2 ranks
# (loader threads, SSL pass, IBM switch, checkpoints; ranks must stay bit-identical).  Output: gpurun_out/r02_mg_*.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_mg_bench_n$N.json 2> gpurun_out/r02_mg_bench_n$N.err; echo "bench N=$N rc=$?"
timeout 300 $TR tools/sweep.py --anet --batches 2 4 8 > gpurun_out/r02_mg_sweep_anet_n$N.txt 2>&1; echo "sweep anet rc=$?"
timeout 400 $TR tools/sweep.py --frames 128 512 1024 --batches 1 8 16 > gpurun_out/r02_mg_sweep_cliplen_n$N.txt 2>&1; echo "sweep clip length rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  tools/train_synthetic.py --videos 6 --epochs 4 --batch 4 --ibm-start 3 --out gpurun_out/train_synth_mg > gpurun_out/r02_mg_train_synth.log 2>&1
echo "train_synthetic (2 ranks) rc=$?"
tail -2 gpurun_out/r02_mg_bench_n$N.json gpurun_out/r02_mg_sweep_anet_n$N.txt gpurun_out/r02_mg_sweep_cliplen_n$N.txt gpurun_out/r02_mg_train_synth.log
