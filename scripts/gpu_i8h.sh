#!/bin/bash
export QEXXC_I8=1
for d in 0 1; do
echo "two-group version, skip-LDTM=$d"
QEXXC_I8_DBG=$d QEXXC_I8_P=1 timeout 300 python scripts/kernel_probe.py 1000 132608 5 2>&1 | grep rowquad
done | tee gpurun_out/i8_ldtm_exp.log
