#!/bin/bash
# wsyrk schedule granularity: time and DRAM traffic per launch at c5 for 16 / 32 / 64 / 128 items per CTA
for it in 16 32 64 128; do
  echo "== QEXXC_WS_ITEMS=$it"
  QEXXC_WS_ITEMS=$it python bench.py --no-configs --steps 3 --warmup 2 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('ms/step', round(d['ms_per_step'],2), 'wsyrk', round(k['wsyrk']['avg_ms'],3), 'rowquad', round(k['rowquad']['avg_ms'],3), 'ws_gb', d['workspace_gb'])"
  QEXXC_WS_ITEMS=$it timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wsyrk_kernel -s 1 -c 1 python scripts/prof_stage.py fwd 1000000 2>&1 | grep -E "dram__bytes|duration"
done 2>&1 | tee gpurun_out/ws_items_sweep.log
