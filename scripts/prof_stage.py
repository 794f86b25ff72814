"""Run one stage of the hot path a few times (for ncu captures):
    python scripts/prof_stage.py ao|fwd|vjp [ngrids] [config] [f64|f32]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qex_b200 import workloads
from qex_b200.engine import XCContext

what = sys.argv[1] if len(sys.argv) > 1 else "ao"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
cfg = sys.argv[3] if len(sys.argv) > 3 else "c5"
prec = sys.argv[4] if len(sys.argv) > 4 else "f64"
wl = workloads.make(cfg, ngrids=G)
ctx = XCContext(nao=wl.nao, ngrids_max=G, ncomp=wl.ncomp, net=workloads.net_spec(wl, prec))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
deriv = 1 if wl.ncomp == 4 else 0
ctx.eval_ao(deriv)
out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
for _ in range(3):
    if what == "ao":
        ctx.eval_ao(deriv)
    elif what == "fwd":
        ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype, out=out, resid=resid)
    else:
        ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
torch.cuda.synchronize()
print("done", what, G)
