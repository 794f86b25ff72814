#!/bin/bash
for v in 0 1; do QEXXC_AO_REG85=$v python scripts/bench_ao.py 1000000 c5 2>&1 | grep -v Warn; QEXXC_AO_REG85=$v python scripts/bench_ao.py 1000000 c5gga 2>&1 | grep -v Warn; QEXXC_AO_REG85=$v python scripts/bench_ao.py 50000 c3 2>&1 | grep -v Warn; done
for v in 0 1; do QEXXC_AO_REG85=$v python scripts/bench_ao.py 1000000 c5 2>&1 | grep -v Warn; done
python -m pytest tests/test_gpu_parity.py -q -k eval_ao 2>&1 | tail -2
QEXXC_AO_REG85=1 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -k "eval_ao or golden" 2>&1 | tail -2
