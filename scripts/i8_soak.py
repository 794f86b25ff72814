"""Soak test of the INT8 contraction pipeline: many back-to-back steps must neither hang nor change a bit.
  python scripts/i8_soak.py [steps] -> 1) `steps` full c5 steps (1e6 points) compared bit for bit with the first one;
  2) 150 re-gridded steps on one context with a different grid size each time, compared with a fresh context."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qex_b200 import workloads
from qex_b200.engine import XCContext

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
wl = workloads.make("c5")
N, G = wl.nao, wl.ngrids
ctx = XCContext(nao=N, ngrids_max=G, ncomp=1, net=workloads.net_spec(wl))
assert ctx.contraction_mode == "int8"
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
dm, th, vb = ctx.dev(wl.dm), ctx.dev(wl.theta), ctx.dev(wl.v_bar)
ref = None
t0 = time.time()
bad = 0
for it in range(steps):
    ctx.eval_ao(0)
    out, resid = ctx.nr_rks_fwd(dm, th, wl.xctype)
    bar = ctx.nr_rks_vjp(th, resid, [wl.e_bar], vb, wl.xctype)
    if ref is None:
        ref = (out.clone(), bar.clone())
    elif it % 10 == 0 or it == steps - 1:
        bad += int(not (torch.equal(out, ref[0]) and torch.equal(bar, ref[1])))
torch.cuda.synchronize()
print(json.dumps({"phase": "c5 steps", "steps": steps, "seconds": round(time.time() - t0, 1), "bitwise_mismatches": bad}), flush=True)
ctx.close()

rng = np.random.default_rng(0)
wl = workloads.make("c5", ngrids=40000)
ctx = XCContext(nao=N, ngrids_max=40000, ncomp=1, net=workloads.net_spec(wl))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env)
worst = 0.0
for it in range(150):
    g = int(rng.integers(1, 40001))
    ctx.set_grid(wl.coords[:g], wl.weights[:g]).eval_ao(0)
    out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
    bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
    if it % 15 == 0:
        c2 = XCContext(nao=N, ngrids_max=g, ncomp=1, net=workloads.net_spec(wl))
        c2.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords[:g], wl.weights[:g]).eval_ao(0)
        o2, r2 = c2.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
        b2 = c2.nr_rks_vjp(wl.theta, r2, [wl.e_bar], wl.v_bar, wl.xctype)
        bad += int(not (torch.equal(out, o2) and torch.equal(bar, b2)))
        c2.close()
torch.cuda.synchronize()
print(json.dumps({"phase": "re-gridded steps", "steps": 150, "bitwise_mismatches_vs_fresh_context": bad}), flush=True)
sys.exit(1 if bad else 0)
