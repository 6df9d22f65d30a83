#!/bin/bash
# compute-sanitizer over the kernels that are new in the second half of round 2: extended GroupNorm, glue kernels, the loss flavours,
# the resident-halo Conv3d_1a kernels, the templated pool backward (through their GPU tests)
set -u
mkdir -p gpurun_out
SAN="compute-sanitizer --report-api-errors no --print-limit 20"
T="tests/test_head_schedule_gpu.py tests/test_msl_gpu.py tests/test_conv1a_u8_gpu.py tests/test_backbone_kernels_gpu.py"
timeout 800 $SAN --tool memcheck --log-file gpurun_out/r02b_memcheck.log python -m pytest $T -q -m gpu > gpurun_out/r02b_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
timeout 1200 $SAN --tool racecheck --log-file gpurun_out/r02b_racecheck.log python -m pytest $T -q -m gpu > gpurun_out/r02b_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
timeout 800 $SAN --tool synccheck --log-file gpurun_out/r02b_synccheck.log python -m pytest $T -q -m gpu > gpurun_out/r02b_synccheck_pytest.log 2>&1; echo "synccheck rc=$?"
for t in memcheck racecheck synccheck; do echo "## $t"; grep -E "SUMMARY|hazard" gpurun_out/r02b_$t.log | tail -3; tail -1 gpurun_out/r02b_${t}_pytest.log; done
