#!/bin/bash
export QEXXC_I8=1
for d in 0 1; do
echo "dbg=$d"
QEXXC_I8_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rowquad_i8_kernel|wsyrk_i8_kernel' -c 4 python scripts/prof_stage.py vjp 262144 c5 2>&1 | grep -E "gpu__time"
done | tee gpurun_out/i8_dbg2.log
