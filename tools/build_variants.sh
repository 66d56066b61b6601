#!/bin/bash
# Build librt_gpu variants for kernel A/B runs: tools/build_variants.sh name:"-DFLAG=.. -DFLAG=.." ...
# Output: build/variants/librt_gpu_<name>.so (build/ is git-ignored but travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
    -Iinclude -Iraytracing-course-hw-public_b200/csrc -o build/variants/librt_gpu_$name.so \
    raytracing-course-hw-public_b200/csrc/rt_gpu.cu -ldl &
done
wait
ls -la build/variants
