#!/bin/bash
export QEXXC_I8=1
for P in 1 2 4; do
echo "P=$P"
QEXXC_I8_P=$P timeout 300 python scripts/kernel_probe.py 1000,2000 132608 5 2>&1 | grep rowquad
done | tee gpurun_out/i8_psplit.log
unset QEXXC_I8_P
timeout 300 python scripts/i8_check.py 131072 c5 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/i8_psplit.log
timeout 300 python scripts/i8_check.py 70000 c5gga 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/i8_psplit.log
