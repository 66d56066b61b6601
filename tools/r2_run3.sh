#!/bin/bash
mkdir -p gpurun_out
P="python tools/bimodal_probe.py --spp 16 --iters 3"
{
$P --tag plain 2>&1 | grep probe
$P --dummy-mb 2 --tag slow 2>&1 | grep probe
} | tee gpurun_out/r2d_bimodal.log
NCU="ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 1 -f"
$NCU -o gpurun_out/r2d_shade_fast $P --tag plain > gpurun_out/r2d_ncu_fast.log 2>&1; echo "ncu fast rc=$?"
$NCU -o gpurun_out/r2d_shade_slow $P --dummy-mb 2 --tag slow > gpurun_out/r2d_ncu_slow.log 2>&1; echo "ncu slow rc=$?"
ls -la gpurun_out/*.ncu-rep
