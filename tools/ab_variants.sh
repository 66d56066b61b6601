#!/bin/bash
# A/B run of librt_gpu variants on the GPU box: parity tests + short bench per variant.
# usage: bash tools/ab_variants.sh <spp> name1 name2 ...
SPP=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  export RT_GPU_LIB=$PWD/build/variants/librt_gpu_$v.so
  [ -z "$AB_SKIP_TESTS" ] && { python -m pytest tests -m gpu -x -q > gpurun_out/ab_${v}_pytest.log 2>&1; echo "$v pytest rc=$? $(tail -1 gpurun_out/ab_${v}_pytest.log)"; }
  python bench.py --steps 2 --warmup 2 --spp $SPP --cpu-budget 0 > gpurun_out/ab_${v}.json 2> gpurun_out/ab_${v}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${v}.json")); k=d["kernel_ms_profiled_step"]
    print("$v value=%.1f extend=%.2f shade=%.2f total=%.2f frac=%.4f"%(d["value"],k["extend"],k["shade"],k["render_total"],d["roofline"]["frac"]))
except Exception as e:
    print("$v failed", e)
PY
done
