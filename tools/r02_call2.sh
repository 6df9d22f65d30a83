#!/bin/bash
# Round-2 GPU call 2: clean per-layer profile (no side-stream overlap) + compute-sanitizer passes over the kernel tests.
set -u
mkdir -p gpurun_out
OTAL_NO_WGRAD_OVERLAP=1 timeout 400 python tools/step_profile.py --top 600 > gpurun_out/r02_step_profile_nooverlap.txt 2>&1; echo "step_profile rc=$?"
SAN="compute-sanitizer --report-api-errors no --print-limit 20"
K="tests/test_conv_gpu.py tests/test_backbone_kernels_gpu.py tests/test_bmp_gpu.py tests/test_msl_gpu.py tests/test_gn_gpu.py tests/test_head_gpu.py tests/test_infer_gpu.py tests/test_conv1a_u8_gpu.py"
OTAL_STAGED=1 timeout 900 $SAN --tool memcheck --log-file gpurun_out/r02_memcheck.log python -m pytest $K -q -m gpu > gpurun_out/r02_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/r02_memcheck.log; tail -2 gpurun_out/r02_memcheck_pytest.log
OTAL_STAGED=1 timeout 900 $SAN --tool racecheck --log-file gpurun_out/r02_racecheck.log python -m pytest $K -q -m gpu > gpurun_out/r02_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/r02_racecheck.log; tail -2 gpurun_out/r02_racecheck_pytest.log
OTAL_STAGED=1 timeout 600 $SAN --tool synccheck --log-file gpurun_out/r02_synccheck.log python -m pytest $K -q -m gpu > gpurun_out/r02_synccheck_pytest.log 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/r02_synccheck.log; tail -2 gpurun_out/r02_synccheck_pytest.log
