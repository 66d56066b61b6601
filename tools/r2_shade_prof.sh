#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2j_pytest.log)"
P="python tools/bimodal_probe.py --spp 32 --iters 2"
$P | grep probe
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 1 -f -o gpurun_out/r2j_shade $P > gpurun_out/r2j_ncu_shade.log 2>&1; echo "ncu rc=$?"
