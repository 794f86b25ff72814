#!/bin/bash
python scripts/bench_ao.py 2>&1 | grep -v Warn | tee gpurun_out/ao_v3.log
python scripts/bench_ao.py 1000000 c5gga 2>&1 | grep -v Warn | tee -a gpurun_out/ao_v3.log
python scripts/bench_ao.py 50000 c3 2>&1 | grep -v Warn | tee -a gpurun_out/ao_v3.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:eval_ao_tiled -s 1 -c 1 python scripts/prof_stage.py ao 262144 2>&1 | grep -E "inst_executed|duration|bank" | tee -a gpurun_out/ao_v3.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_gputests_b.log
python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 600 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
