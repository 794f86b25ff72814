#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -k "n_rank" -x -q 2>&1 | tail -4 | tee gpurun_out/n2b_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02c_bench_n2.json').read().strip().splitlines() if l.startswith("{")][-1])
print(2, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e'].get('bitwise_equal_to_device_path'), d['parity_vs_n1']['ok'])
PY
timeout 600 python bench.py --no-configs --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02c_bench_n1.json').read().strip().splitlines() if l.startswith("{")][-1])
print(1, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e'].get('bitwise_equal_to_device_path'))
PY
