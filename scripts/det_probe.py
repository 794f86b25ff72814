"""Run-to-run bitwise determinism probe of the hot path:  python scripts/det_probe.py [ngrids] [config]
Evaluates AO, rho and three full fwd+VJP steps on the same inputs and reports every output that is not
bit-identical between runs (a data race shows up here long before it shows up in a parity tolerance)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from qex_b200 import workloads  # noqa: E402
from qex_b200.engine import XCContext  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "c5"
wl = workloads.make(cfg, ngrids=G)
ctx = XCContext(nao=wl.nao, ngrids_max=G, ncomp=wl.ncomp, net=workloads.net_spec(wl))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env)
deriv = 1 if wl.ncomp == 4 else 0
ctx.set_grid(wl.coords, wl.weights).eval_ao(deriv)
ao1 = ctx.get_ao(wl.ncomp).clone()
ctx.eval_ao(deriv)
ok = torch.equal(ao1, ctx.get_ao(wl.ncomp))
del ao1
rho1 = ctx.eval_rho(wl.dm, wl.ncomp, 1).clone()
ok &= torch.equal(rho1, ctx.eval_rho(wl.dm, wl.ncomp, 1))
outs = []
for i in range(3):
    out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
    bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
    outs.append((out.clone(), bar.clone(), resid.clone()))
for i in (1, 2):
    for k, name in enumerate(("out", "bar", "resid")):
        a, b = outs[0][k], outs[i][k]
        if not torch.equal(a, b):
            ok = False
            d = (a - b).abs()
            print("MISMATCH run", i, name, "max", d.max().item(), "count", int((d > 0).sum()))
print(cfg, G, "deterministic" if ok else "NOT DETERMINISTIC")
sys.exit(0 if ok else 1)
