#!/bin/bash
scripts/probe/i8_rate.bin 2>&1 | tee gpurun_out/i8_rate.log
timeout 900 python -m pytest tests/test_scf_masked.py tests/test_gpu_mlp_wide.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/masked_tests.log
python bench.py --config c5w512 --no-configs --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5w512', d['ms_per_step'], {k:(v['avg_ms'],v['launches']) for k,v in d['roofline']['kernels'].items()})" | tee gpurun_out/c5w512.log
