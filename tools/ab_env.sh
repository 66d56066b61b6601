#!/bin/bash
# A/B run of librt_gpu variants with per-run environment: tools/ab_env.sh <spp> "name[:VAR=val[,VAR=val]]" ...
SPP=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  [ "$spec" != "$v" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  tag=$(echo "$spec" | tr ':=,' '___')
  env $envs RT_GPU_LIB=$PWD/build/variants/librt_gpu_$v.so python bench.py --steps 2 --warmup 2 --spp $SPP --cpu-budget 0 > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${tag}.json")); k=d["kernel_ms_profiled_step"]
    print("$spec value=%.1f extend=%.2f light=%.2f shade=%.2f total=%.2f frac=%.4f"%(d["value"],k["extend"],k.get("lightpdf",0),k["shade"],k["render_total"],d["roofline"]["frac"]))
except Exception as e:
    print("$spec failed", e)
PY
done
