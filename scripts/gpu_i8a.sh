#!/bin/bash
# ncu capture of the INT8 rowquad kernel (c5, 131072 points)
set -x
export QEXXC_I8=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowquad_i8 -s 1 -c 1 -o gpurun_out/rq_i8_a python scripts/prof_stage.py fwd 131072 c5 > gpurun_out/rq_i8_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:slice_ao -c 1 -o gpurun_out/slice_ao_a python scripts/prof_stage.py fwd 131072 c5 >> gpurun_out/rq_i8_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
