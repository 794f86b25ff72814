#!/bin/bash
set -x
for cfg in c5 c5gga; do
  timeout 300 python scripts/i8_check.py 131072 $cfg 2>&1 | tail -3 | tee -a gpurun_out/i8_check2.log
done
timeout 300 python scripts/i8_check.py 100000 c5 2>&1 | tail -3 | tee -a gpurun_out/i8_check2.log
export QEXXC_I8=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'i8|slice|colmax|blk_exp' -c 40 python scripts/prof_stage.py vjp 131072 c5 2>&1 | grep -E "^  [a-z_]+.*\(|gpu__time" | paste - - | awk '{print $1, $(NF-1), $NF}' | tee gpurun_out/i8_launches.log
