#!/bin/bash
# round 2, GPU call 1: parity tests, bench, then the bimodal probe (separate processes)
mkdir -p gpurun_out
bash tools/gpu_cycle.sh r2a 1000 0
for i in 1 2 3 4 5 6; do python tools/bimodal_probe.py --tag plain$i 2>&1 | grep probe; done | tee gpurun_out/r2a_bimodal.log
for mb in 2 34 514 1026 4098 20000; do python tools/bimodal_probe.py --dummy-mb $mb --tag dummy 2>&1 | grep probe; done | tee -a gpurun_out/r2a_bimodal.log
