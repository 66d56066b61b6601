#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2e_pytest.log)"
{
for v in "" r1 pad f4 m7 c256 c64; do
  if [ -n "$v" ]; then export RT_GPU_LIB=$PWD/build/variants/librt_gpu_$v.so; else unset RT_GPU_LIB; fi
  python tools/bimodal_probe.py --tag "v=${v:-default} plain" 2>&1 | grep probe
  python tools/bimodal_probe.py --dummy-mb 2 --tag "v=${v:-default} slow" 2>&1 | grep probe
done
} | tee gpurun_out/r2e_variants.log
