#!/bin/bash
# 8-GPU session (INT8 contractions default): N-rank vs 1-rank parity (2, 4, 8 ranks), strong-scaling bench at N = 8, 4, 2, reference arm at N = 2
timeout 900 python -m pytest tests/test_gpu_round2.py -k "n_rank" -x -q 2>&1 | tail -6 | tee gpurun_out/n8b_parity.log
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02b_bench_n$n.json 2> gpurun_out/r02b_bench_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02b_bench_n$n.json').read().strip().splitlines()[-1])
print($n, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['collectives'], d['parity_vs_n1'])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02b_bench_ref_n2.json 2> gpurun_out/r02b_bench_ref_n2.err
tail -c 600 gpurun_out/r02b_bench_ref_n2.json
