#!/bin/bash
# 8-GPU session: N-rank vs 1-rank parity (2, 4, 8 ranks), strong-scaling bench at N = 4 and 8
timeout 900 python -m pytest tests/test_gpu_round2.py -k "n_rank" -x -q 2>&1 | tail -6 | tee gpurun_out/n8_parity.log
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$n.json').read().strip().splitlines()[-1])
print($n, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['collectives'], d['parity_vs_n1']['ok'])
PY
done
