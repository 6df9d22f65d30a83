#!/bin/bash
# First GPU call of round 2 (run through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_first_gpu_call.sh'
# 1. the default GPU suite (code written after round 1's GPU budget ran out: Trainer graph cache / fixed target slots,
#    per-device kernel configuration, dataset / config modules),
# 2. the STAGED raw-uint8 Conv3d_1a kernels: opt-in tests, micro-benchmark, A/B of the whole step,
# 3. ncu of the two conv1a forms, 4. the synthetic-dataset training run.
# Everything lands in gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest default rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log
OTAL_STAGED=1 python -m pytest tests/test_conv1a_u8_gpu.py tests/test_model_anet_gpu.py -q > gpurun_out/r02_pytest_staged.log 2>&1; echo "pytest staged rc=$?" | tee -a gpurun_out/r02_pytest_staged.log
timeout 300 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench.txt 2>&1; echo "conv1a_bench rc=$?"
for flag in 0 1 0 1; do
  OTAL_U8_CONV1A=$flag timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_u8_${flag}_$RANDOM.json 2> gpurun_out/r02_bench_err.log
  echo "bench OTAL_U8_CONV1A=$flag rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|conv_wgrad|border_class|clip_ingest' -c 12 \
  -o gpurun_out/r02_conv1a python tools/conv1a_bench.py --ncu > gpurun_out/r02_conv1a_ncu.log 2>&1; echo "ncu rc=$?"
# pipeline-depth switch for the 2-stage shapes (3 stages + 1 staging buffer instead of 2 + 2)
for flag in 0 1 0 1; do
  if [ $flag = 1 ]; then export OTAL_CONV_PREFER_STAGES=1; else unset OTAL_CONV_PREFER_STAGES; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_stages_${flag}_$RANDOM.json 2>> gpurun_out/r02_bench_err.log
  echo "bench OTAL_CONV_PREFER_STAGES=$flag rc=$?"
done
unset OTAL_CONV_PREFER_STAGES
# 64-wide N blocks for the large 1x1 convs (more loads in flight per SM)
for flag in 0 1 0 1; do
  if [ $flag = 1 ]; then export OTAL_CONV_1X1_BN64=1; else unset OTAL_CONV_1X1_BN64; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bn64_${flag}_$RANDOM.json 2>> gpurun_out/r02_bench_err.log
  echo "bench OTAL_CONV_1X1_BN64=$flag rc=$?"
done
unset OTAL_CONV_1X1_BN64
# both launch-shape switches together: 64-wide N blocks + one staging buffer = 4 stages (128 KB of A) in flight per SM
for flag in 0 1 0 1; do
  if [ $flag = 1 ]; then export OTAL_CONV_1X1_BN64=1 OTAL_CONV_PREFER_STAGES=1; else unset OTAL_CONV_1X1_BN64 OTAL_CONV_PREFER_STAGES; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bn64stages_${flag}_$RANDOM.json 2>> gpurun_out/r02_bench_err.log
  echo "bench BN64+PREFER_STAGES=$flag rc=$?"
done
unset OTAL_CONV_1X1_BN64 OTAL_CONV_PREFER_STAGES
# b1a + b2a of every inception block as one forward and one weight-gradient launch (host-side fusion, 18 launches fewer)
for flag in 0 1 0 1; do
  if [ $flag = 1 ]; then export OTAL_FUSE_B12A=1; else unset OTAL_FUSE_B12A; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_fuse12_${flag}_$RANDOM.json 2>> gpurun_out/r02_bench_err.log
  echo "bench OTAL_FUSE_B12A=$flag rc=$?"
done
unset OTAL_FUSE_B12A
# where the step goes now, per layer (event-timed eager pass), and a source-level look at the HBM-bound 1x1 convs, which run
# at ~22 % of the copy roofline (profiles/r01_ncu_full_kernels_summary_v2.txt ids 8-10: Mixed_3c.b0 fwd / dgrad / wgrad)
timeout 600 python tools/step_profile.py > gpurun_out/r02_step_profile.txt 2>&1; echo "step_profile rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm' -s 5 -c 2 \
  -o gpurun_out/r02_conv1x1 python tools/ncu_targets.py > gpurun_out/r02_conv1x1_ncu.log 2>&1; echo "ncu 1x1 rc=$?"
timeout 900 python tools/train_synthetic.py --videos 6 --epochs 12 --batch 4 --ibm-start 3 --out gpurun_out/train_synth > gpurun_out/r02_train_synth.log 2>&1
echo "train_synthetic rc=$?"
tail -3 gpurun_out/r02_pytest_gpu.log gpurun_out/r02_pytest_staged.log gpurun_out/r02_conv1a_bench.txt gpurun_out/r02_train_synth.log
