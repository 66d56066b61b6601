#!/bin/bash
# strong-scaling bench lines (config 4) at N GPUs of this box: bash tools/r2_scale.sh N [tag]
N=${1:-8}; TAG=${2:-r2_v3}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
fi
echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","scaling","n_gpus","parity_check","kernel_ms_profiled_step")}, "e2e", d["e2e"]["value"])
PY
