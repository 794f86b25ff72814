#!/bin/bash
scripts/probe/tc_probe16.bin > gpurun_out/tc_probe16.log 2>&1; cat gpurun_out/tc_probe16.log
timeout 600 python -m pytest tests/test_gpu_mlp_wide.py -x -q 2>&1 | tail -15 | tee gpurun_out/wide_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_ao_tiled -s 1 -c 1 -o gpurun_out/ao_r02 python scripts/prof_stage.py ao 262144 > gpurun_out/ncu_ao.log 2>&1; tail -2 gpurun_out/ncu_ao.log
