#!/bin/bash
# final single-GPU profiling session of round 2
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_gputests_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_ref_final.json
# launch list of the bench command (shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/bench_under_ncu.log 2>&1
# full captures: FP32 tensor-core MLP kernels, the wide-network GEMM, AO evaluator (GGA)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_fwd -s 1 -c 1 -o gpurun_out/tc_fwd_r02 python scripts/prof_stage.py fwd 262144 c5 f32 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_vjp -s 1 -c 1 -o gpurun_out/tc_vjp_r02 python scripts/prof_stage.py vjp 262144 c5 f32 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm64 -s 4 -c 3 -o gpurun_out/gemm64_r02 python scripts/prof_stage.py vjp 32768 c5w512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_ao_tiled -s 1 -c 1 -o gpurun_out/ao_gga_r02 python scripts/prof_stage.py ao 262144 c5gga > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
