#!/bin/bash
# INT8 contraction profiles: launch list of the bench command, full captures of the four new kernels, DRAM traffic at the full c5 size
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_under_ncu_r02b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowquad_i8_kernel -s 1 -c 1 -o gpurun_out/rq_i8_r02b python scripts/prof_stage.py fwd 262144 c5 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wsyrk_i8_kernel -s 1 -c 1 -o gpurun_out/ws_i8_r02b python scripts/prof_stage.py fwd 262144 c5 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:slice_rows -c 1 -o gpurun_out/slice_rows_r02b python scripts/prof_stage.py fwd 262144 c5 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:slice_cols -s 1 -c 1 -o gpurun_out/slice_cols_r02b python scripts/prof_stage.py fwd 262144 c5 > /dev/null 2>&1
# DRAM traffic of the two contraction kernels at the FULL c5 size (1e6 points)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'rowquad_i8_kernel|wsyrk_i8_kernel' -c 4 python scripts/prof_stage.py vjp 1000000 c5 2>&1 | grep -E "i8_kernel|dram__|gpu__time" > gpurun_out/traffic_i8_c5.log
ls -la gpurun_out/*r02b*
