#!/bin/bash
# multi-GPU evidence: un-skipped tests/test_gpu_multi.py + strong-scaling bench with parity_check.  usage: bash tools/r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2_multi${N}_gpus.log
python -m pytest tests/test_gpu_multi.py -m gpu -v -rs > gpurun_out/r2_multi${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_multi${N}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 \
  > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","scaling","n_gpus","parity_check","kernel_ms_profiled_step")}, d["e2e"]["value"])
PY
if [ "$2" = "all" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling weak \
    > gpurun_out/r2_bench_${N}gpu_weak.json 2> gpurun_out/r2_bench_${N}gpu_weak.err; echo "weak bench rc=$?"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --config c5 \
    > gpurun_out/r2_bench_${N}gpu_c5.json 2> gpurun_out/r2_bench_${N}gpu_c5.err; echo "c5 bench rc=$?"
  python - <<PY
import json
for f in ("weak","c5"):
    d=json.load(open("gpurun_out/r2_bench_${N}gpu_%s.json"%f))
    print(f, {k:d[k] for k in ("value","ms_per_step","scaling","n_gpus","parity_check")}, d["e2e"]["value"], d["config"]["workload"])
PY
fi
