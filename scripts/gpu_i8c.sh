#!/bin/bash
set -x
export QEXXC_I8=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wsyrk_i8_kernel -s 1 -c 1 -o gpurun_out/ws_i8_a python scripts/prof_stage.py fwd 131072 c5 > gpurun_out/ws_i8_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
