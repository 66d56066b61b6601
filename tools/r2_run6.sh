#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2h_pytest.log)"
grep -h "primary ids\|1000x1000x2\|device tree\|sah build" gpurun_out/r2h_pytest.log
