#!/bin/bash
# 2-GPU session: N-rank vs 1-rank parity on hardware, then the strong-scaling bench line + reference arm at N = 2
timeout 900 python -m pytest tests/test_gpu_round2.py -k "n_rank" -x -q 2>&1 | tail -6 | tee gpurun_out/n2_parity.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 1500 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err
tail -c 800 gpurun_out/r02_bench_ref_n2.json
