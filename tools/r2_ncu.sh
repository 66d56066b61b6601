#!/bin/bash
# Round-2 evidence run (1 GPU): reference arm, full bench, ncu launch list, `ncu --set full` of k_extend / k_shade at
# 128 spp (128 Mi paths per batch), DRAM bytes of the k_extend launches, launch list of one scene upload (device build).
TAG=${1:-r2_v3}
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; RC=$?
echo "bench rc=$RC"; head -c 400 gpurun_out/${TAG}_bench.json; echo
[ $RC -ne 0 ] && exit $RC
CMD="python bench.py --steps 1 --warmup 1 --spp 16 --cpu-budget 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_extend -c 8 --csv \
    --log-file gpurun_out/${TAG}_extend_dram.csv $CMD > gpurun_out/${TAG}_ncu_extend_dram.log 2>&1
echo "ncu extend dram rc=$?"
CMD2="python tools/bimodal_probe.py --spp 128 --iters 2"
$CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 1 -c 2 -f -o gpurun_out/${TAG}_extend $CMD2 > gpurun_out/${TAG}_ncu_extend.log 2>&1
echo "ncu extend rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -f -o gpurun_out/${TAG}_shade $CMD2 > gpurun_out/${TAG}_ncu_shade.log 2>&1
echo "ncu shade rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_lightpdf -s 0 -c 2 -f -o gpurun_out/${TAG}_lightpdf $CMD2 > gpurun_out/${TAG}_ncu_lightpdf.log 2>&1
echo "ncu lightpdf rc=$?"
ls -la gpurun_out/${TAG}*
exit 0
