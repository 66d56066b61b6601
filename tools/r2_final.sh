#!/bin/bash
# Final evidence of the round on one GPU: full GPU test log (with the printed parity figures) + tools/r2_ncu.sh
TAG=${1:-r2_v6}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_gpu_pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gpu_pytest_1gpu.log
bash tools/r2_ncu.sh $TAG
