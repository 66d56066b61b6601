#!/bin/bash
# A/B run of runtime knobs: bash tools/ab_env.sh <spp> "NAME=VAL ..." "NAME=VAL ..." ...
SPP=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 2 --warmup 2 --spp $SPP --cpu-budget 0 > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abenv_$i.json")); k=d["kernel_ms_profiled_step"]
    print("[$envs] value=%.1f extend=%.2f shade=%.2f sort=%.2f total=%.2f frac=%.4f"%(d["value"],k["extend"],k["shade"],k.get("sort",0),k["render_total"],d["roofline"]["frac"]))
except Exception as e:
    print("[$envs] failed", e)
PY
done
