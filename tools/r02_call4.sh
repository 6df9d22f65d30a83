#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/conv_timeline.py > gpurun_out/r02_conv_timeline.txt 2>&1; echo "timeline rc=$?"; tail -5 gpurun_out/r02_conv_timeline.txt
