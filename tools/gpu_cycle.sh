#!/bin/bash
# One GPU iteration: parity tests, a short bench, then (only if both exited 0) an ncu capture of k_extend / k_shade.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [spp] [ncu: 0|1]
TAG=${1:-run}; SPP=${2:-100}; NCU=${3:-1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; RC=$?
echo "pytest rc=$RC"; tail -3 gpurun_out/${TAG}_pytest.log
[ $RC -ne 0 ] && exit $RC
python bench.py --steps 2 --warmup 3 --spp $SPP --cpu-budget 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; RC=$?
echo "bench rc=$RC"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print(d["value"], d["mrays_per_s"], d["kernel_ms_profiled_step"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e:
    print("no bench json", e)
PY
[ $RC -ne 0 ] && exit $RC
if [ "$NCU" = "1" ]; then
  python bench.py --steps 1 --warmup 1 --spp 8 --cpu-budget 0 > gpurun_out/${TAG}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_extend -s 1 -c 2 -f -o gpurun_out/${TAG}_extend \
      python bench.py --steps 1 --warmup 1 --spp 8 --cpu-budget 0 > gpurun_out/${TAG}_ncu_extend.log 2>&1
  echo "ncu extend rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -f -o gpurun_out/${TAG}_shade \
      python bench.py --steps 1 --warmup 1 --spp 8 --cpu-budget 0 > gpurun_out/${TAG}_ncu_shade.log 2>&1
  echo "ncu shade rc=$?"
fi
exit 0
