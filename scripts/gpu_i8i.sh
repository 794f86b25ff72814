#!/bin/bash
export QEXXC_I8=1
timeout 300 python scripts/i8_check.py 131072 c5 2>&1 | tail -1 | cut -c1-420 | tee gpurun_out/i8_amn.log
timeout 300 python scripts/i8_check.py 70000 c5gga 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/i8_amn.log
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -5 | tee -a gpurun_out/i8_amn.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'i8|slice|colmax|blk_exp' -c 12 python scripts/prof_stage.py vjp 131072 c5 2>&1 | grep -E "^  [a-z_]+.*\(|gpu__time" | paste - - | awk '{print $1, $(NF-1), $NF}' | tee -a gpurun_out/i8_amn.log
