#!/bin/bash
# one-off GPU session: tcgen05 descriptor probe, AO parity, AO (P, threads) sweep
scripts/probe/tc_probe.bin > gpurun_out/tc_probe.log 2>&1; cat gpurun_out/tc_probe.log
python -m pytest tests/test_gpu_parity.py -x -q -k "eval_ao or golden" 2>&1 | tail -3
python -m pytest tests/test_golden.py tests/test_zz_pyscf_pin.py -q -m gpu 2>&1 | tail -3
(python scripts/bench_ao.py
 for cfg in "12 512" "8 512" "16 512" "4 256" "8 256" "12 384" "8 1024" "12 1024"; do set -- $cfg; QEXXC_AO_P=$1 QEXXC_AO_THREADS=$2 python scripts/bench_ao.py; done
 python scripts/bench_ao.py 1000000 c5gga
 QEXXC_AO_P=4 QEXXC_AO_THREADS=512 python scripts/bench_ao.py 1000000 c5gga
 python scripts/bench_ao.py 50000 c3 ) 2>&1 | grep -v Warn | tee gpurun_out/ao_sweep.log
timeout 300 python -m pytest tests/test_gpu_mlp_tc.py -x -q 2>&1 | tail -15 | tee gpurun_out/tc_tests.log
