#!/bin/bash
export QEXXC_I8=1
for v in 0 1; do
if [ $v = 1 ]; then export QEXXC_I8_LD128=1; echo "128-bit loads"; else unset QEXXC_I8_LD128; echo "256-bit no-allocate loads"; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rowquad_i8_kernel' -c 3 python scripts/prof_stage.py vjp 1000000 c5 2>&1 | grep -E "gpu__time"
timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from qex_b200 import workloads
from qex_b200.engine import XCContext
G=1000000
wl = workloads.make("c5", ngrids=G)
ctx = XCContext(nao=wl.nao, ngrids_max=G, ncomp=1, net=workloads.net_spec(wl))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
ctx.eval_ao(0)
out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
torch.cuda.synchronize()
ctx.profile_enable(True)
for _ in range(5):
    ctx.eval_ao(0)
    out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
    bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
torch.cuda.synchronize()
pr = ctx.profile_read()
print({k: round(v[0]/max(v[1],1),3) for k,v in pr.items()})
PY
done 2>&1 | tee gpurun_out/i8_ld_ab.log
