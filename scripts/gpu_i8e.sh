#!/bin/bash
set -x
timeout 600 python bench.py --no-configs --steps 5 --warmup 3 > gpurun_out/bench_i8_a.json 2> gpurun_out/bench_i8_a.err; tail -c 600 gpurun_out/bench_i8_a.err
