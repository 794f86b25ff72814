#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_mlp_tc.py -x -q 2>&1 | tail -25 | tee gpurun_out/tc_tests.log
