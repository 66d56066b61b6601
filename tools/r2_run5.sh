#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2g_pytest.log)"
RT_TIMING=1 python tools/upload_timing.py 2>&1 | tail -12
python bench.py --steps 3 --warmup 3 --cpu-budget 2 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench.json"))
print(d["value"], d["e2e"], d["kernel_ms_profiled_step"], d["roofline"]["frac"], d["host_s"])
PY
