#!/bin/bash
# two-stream queue halves (default) against RT_NO_SPLIT=1, at the 8-GPU strong-scaling share (125 spp) and at 1000 spp
mkdir -p gpurun_out
for spp in 125 1000; do for env in "RT_NO_SPLIT=1" "RT_X=0" "RT_NO_SPLIT=1" "RT_X=0"; do
  env $env python bench.py --steps 3 --warmup 3 --spp $spp --cpu-budget 0 > gpurun_out/ab_split_tmp.json 2> gpurun_out/ab_split_tmp.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_split_tmp.json"))
    print("spp $spp $env value=%.1f ms_per_step=%.2f e2e=%.1f launches=%d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
except Exception as e:
    print("spp $spp $env failed", e)
PY
done; done
