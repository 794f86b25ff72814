#!/bin/bash
export QEXXC_I8=1
for t in 0 1; do
if [ $t = 1 ]; then export QEXXC_I8_T=1; echo "wsyrk A = column-scaled planes, K-major"; else unset QEXXC_I8_T; echo "wsyrk A = row-scaled planes, MN-major"; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'wsyrk_i8_kernel' -c 3 python scripts/prof_stage.py fwd 262144 c5 2>&1 | grep -E "gpu__time"
done 2>&1 | tee gpurun_out/i8_amajor.log
