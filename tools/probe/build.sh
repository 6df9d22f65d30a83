#!/bin/bash
# builds the developer probes into tools/probe/bin (git-ignored; travels to the GPU box with gpurun)
set -e
cd "$(dirname "$0")"; mkdir -p bin
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -o bin/${f%.cu} $f ../../opental_b200/csrc/api.cu -lcudart
done
