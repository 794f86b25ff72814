#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_hostapi.py tests/test_gpu_round2.py -m gpu -x -q -k "full_size or determin or c5" 2>&1 | tail -5
python bench.py --no-configs --steps 3 --warmup 2 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('ms/step', round(d['ms_per_step'],2), {a:round(v['avg_ms'],3) for a,v in k.items()})"
QEXXC_NO_PAIR=1 python bench.py --no-configs --steps 3 --warmup 2 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('NO_PAIR ms/step', round(d['ms_per_step'],2), {a:round(v['avg_ms'],3) for a,v in k.items()})"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"rowquad_kernel|wsyrk_kernel" -s 2 -c 2 python scripts/prof_stage.py fwd 1000000 2>&1 | grep -E "rowquad_kernel|wsyrk_kernel|dram__bytes|duration"
python bench.py --config c5gga --no-configs --steps 2 --warmup 2 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('c5gga ms/step', round(d['ms_per_step'],2), {a:round(v['avg_ms'],3) for a,v in k.items()})"
for v in 0 1; do QEXXC_AO_REG85=$v python scripts/bench_ao.py 1000000 c5 2>&1 | grep -v Warn; QEXXC_AO_REG85=$v python scripts/bench_ao.py 1000000 c5gga 2>&1 | grep -v Warn; done
python -m pytest tests/test_gpu_parity.py -q -k eval_ao 2>&1 | tail -2
QEXXC_AO_REG85=1 python -m pytest tests/test_gpu_parity.py -q -k eval_ao 2>&1 | tail -2
