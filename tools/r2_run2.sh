#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,serial,uuid,memory.used,memory.total,pci.bus_id --format=csv | tee gpurun_out/r2c_bimodal.log
P="python tools/bimodal_probe.py"
{
$P --tag plainA 2>&1 | grep probe
RT_TIMING=1 $P --tag plainTIMING 2>&1 | grep "probe\|arena"
$P --tag plainB 2>&1 | grep probe
$P --dummy-mb 2 --tag d 2>&1 | grep probe
$P --dummy-mb 1026 --tag d 2>&1 | grep probe
$P --dummy-mb 4098 --tag d 2>&1 | grep probe
$P --tag plainC 2>&1 | grep probe
$P --paths 33554432 --tag paths32M 2>&1 | grep probe
$P --paths 33554432 --dummy-mb 2 --tag paths32M 2>&1 | grep probe
} | tee -a gpurun_out/r2c_bimodal.log
