#!/bin/bash
# Round-end style measurement on the GPU box: parity tests, reference arm, full bench (config 4), ncu launch list
# and one `ncu --set full` capture of k_extend / k_shade.   usage: bash tools/gpu_full.sh <tag>
TAG=${1:-full}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; RC=$?
echo "pytest rc=$RC"; tail -3 gpurun_out/${TAG}_pytest.log
[ $RC -ne 0 ] && exit $RC
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; RC=$?
echo "bench rc=$RC"; head -c 300 gpurun_out/${TAG}_bench.json; echo
[ $RC -ne 0 ] && exit $RC
CMD="python bench.py --steps 1 --warmup 1 --spp 16 --cpu-budget 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 1 -c 3 -f -o gpurun_out/${TAG}_extend $CMD > gpurun_out/${TAG}_ncu_extend.log 2>&1
echo "ncu extend rc=$?"
# DRAM bytes of ALL k_extend launches of one step (16 spp = one batch = 8 launches) -> bytes per extension ray
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_extend -c 8 --csv \
    --log-file gpurun_out/${TAG}_extend_dram.csv $CMD > gpurun_out/${TAG}_ncu_extend_dram.log 2>&1
echo "ncu extend dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -f -o gpurun_out/${TAG}_shade $CMD > gpurun_out/${TAG}_ncu_shade.log 2>&1
echo "ncu shade rc=$?"
exit 0
