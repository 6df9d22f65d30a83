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
OTAL_STAGED=1 timeout 300 python -m pytest tests/test_conv1a_u8_gpu.py tests/test_model_anet_gpu.py -q > gpurun_out/r02_pytest_staged.log 2>&1; echo "pytest staged rc=$?" | tee -a gpurun_out/r02_pytest_staged.log
timeout 300 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench.txt 2>&1; echo "conv1a_bench rc=$?"
# A/B of every staged switch: two interleaved passes over (baseline, each switch alone, the two launch-shape switches together);
# `value` of each JSON line is the CUDA-event step time with resident inputs (no e2e leg, no CPU baseline: ~40 s per run)
for pass in 1 2; do
  for cfg in base OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 "OTAL_CONV_1X1_BN64 OTAL_CONV_PREFER_STAGES" OTAL_FUSE_B12A OTAL_CONV_KSPLIT; do
    unset OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 OTAL_FUSE_B12A OTAL_CONV_KSPLIT
    if [ "$cfg" != base ]; then for v in $cfg; do export $v=1; done; fi
    tag=$(echo "$cfg" | tr ' ' '+')
    timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ab_${tag}_pass${pass}.json 2>> gpurun_out/r02_bench_err.log
    echo "bench [$tag] pass $pass rc=$? $(python -c "import json,sys; print(json.load(open('gpurun_out/r02_ab_${tag}_pass${pass}.json'))['ms_per_step'])" 2>/dev/null) ms/step"
  done
done
unset OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 OTAL_FUSE_B12A OTAL_CONV_KSPLIT
python tools/ab_summary.py gpurun_out > gpurun_out/r02_ab_summary.txt 2>&1; cat gpurun_out/r02_ab_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|conv_wgrad|border_class|clip_ingest' -c 12 \
  -o gpurun_out/r02_conv1a python tools/conv1a_bench.py --ncu > gpurun_out/r02_conv1a_ncu.log 2>&1; echo "ncu conv1a rc=$?"
# where the step goes now, per layer (event-timed eager pass), and a source-level look at the HBM-bound 1x1 convs, which run
# at ~22 % of the copy roofline (profiles/r01_ncu_full_kernels_summary_v2.txt ids 8-10: Mixed_3c.b0 fwd / dgrad / wgrad)
timeout 600 python tools/step_profile.py > gpurun_out/r02_step_profile.txt 2>&1; echo "step_profile rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm' -s 5 -c 2 \
  -o gpurun_out/r02_conv1x1 python tools/ncu_targets.py > gpurun_out/r02_conv1x1_ncu.log 2>&1; echo "ncu 1x1 rc=$?"
timeout 900 python tools/train_synthetic.py --videos 6 --epochs 12 --batch 4 --ibm-start 3 --out gpurun_out/train_synth > gpurun_out/r02_train_synth.log 2>&1
echo "train_synthetic rc=$?"
tail -3 gpurun_out/r02_pytest_gpu.log gpurun_out/r02_pytest_staged.log gpurun_out/r02_conv1a_bench.txt gpurun_out/r02_train_synth.log
