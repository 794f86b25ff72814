#!/bin/bash
for a in "1000000 c5" "1000000 c5gga" "50000 c3"; do python scripts/bench_ao.py $a 2>&1 | grep -v Warn; done | tee gpurun_out/ao_v4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_ao_tiled -s 1 -c 1 -o gpurun_out/ao_r02b python scripts/prof_stage.py ao 262144 > gpurun_out/ncu_ao.log 2>&1; tail -1 gpurun_out/ncu_ao.log
timeout 900 python -m pytest tests/test_scf_masked.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/masked_tests.log
