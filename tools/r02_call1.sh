#!/bin/bash
# Round-2 GPU call 1: default suite, staged tests, conv1a micro-benchmark, one A/B pass over the staged switches,
# ncu of conv1a / 1x1, step profile, short synthetic training run, first sanitizer pass.  Output: gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest default rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log
OTAL_STAGED=1 timeout 300 python -m pytest tests/test_conv1a_u8_gpu.py tests/test_model_anet_gpu.py -q > gpurun_out/r02_pytest_staged.log 2>&1; echo "pytest staged rc=$?" | tee -a gpurun_out/r02_pytest_staged.log
timeout 300 python tools/conv1a_bench.py > gpurun_out/r02_conv1a_bench.txt 2>&1; echo "conv1a_bench rc=$?"
for cfg in base OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 "OTAL_CONV_1X1_BN64 OTAL_CONV_PREFER_STAGES" OTAL_FUSE_B12A OTAL_CONV_KSPLIT base2; do
  unset OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 OTAL_FUSE_B12A OTAL_CONV_KSPLIT
  if [ "$cfg" != base ] && [ "$cfg" != base2 ]; then for v in $cfg; do export $v=1; done; fi
  tag=$(echo "$cfg" | tr ' ' '+')
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ab_${tag}_pass1.json 2>> gpurun_out/r02_bench_err.log
  echo "bench [$tag] rc=$? $(python -c "import json,sys; print(json.load(open('gpurun_out/r02_ab_${tag}_pass1.json'))['ms_per_step'])" 2>/dev/null) ms/step"
done
unset OTAL_U8_CONV1A OTAL_CONV_PREFER_STAGES OTAL_CONV_1X1_BN64 OTAL_FUSE_B12A OTAL_CONV_KSPLIT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|conv_wgrad|border_class|clip_ingest' -c 12 \
  -o gpurun_out/r02_conv1a python tools/conv1a_bench.py --ncu > gpurun_out/r02_conv1a_ncu.log 2>&1; echo "ncu conv1a rc=$?"
timeout 400 python tools/step_profile.py > gpurun_out/r02_step_profile.txt 2>&1; echo "step_profile rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm' -s 5 -c 2 \
  -o gpurun_out/r02_conv1x1 python tools/ncu_targets.py > gpurun_out/r02_conv1x1_ncu.log 2>&1; echo "ncu 1x1 rc=$?"
timeout 300 python tools/train_synthetic.py --videos 4 --epochs 6 --batch 4 --ibm-start 2 --out gpurun_out/train_synth > gpurun_out/r02_train_synth.log 2>&1
echo "train_synthetic rc=$?"
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_a.log python -m pytest tests/test_bmp_gpu.py tests/test_msl_gpu.py tests/test_conv_gpu.py -q -x -m gpu > gpurun_out/r02_memcheck_a_pytest.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02_memcheck_a.log
tail -3 gpurun_out/r02_pytest_gpu.log gpurun_out/r02_pytest_staged.log gpurun_out/r02_conv1a_bench.txt gpurun_out/r02_train_synth.log
