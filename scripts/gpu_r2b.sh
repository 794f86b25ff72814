#!/bin/bash
# round-2 second half: full GPU suite + the driver's bench command + reference arm
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02b_gputests.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 400 gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.json
