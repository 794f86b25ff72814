"""rowquad on the INT8 tensor cores (QEXXC_I8=1) against the FP64 DMMA path: max-norm difference of rho and of the
fwd+VJP outputs, and kernel times.   python scripts/i8_check.py [ngrids] [c5|c5gga]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qex_b200 import workloads
from qex_b200.engine import XCContext

G = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
cfg = sys.argv[2] if len(sys.argv) > 2 else "c5"
wl = workloads.make(cfg, ngrids=G)
N = wl.nao
ctx = XCContext(nao=N, ngrids_max=G, ncomp=wl.ncomp, net=workloads.net_spec(wl))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
deriv = 1 if wl.ncomp == 4 else 0
ctx.eval_ao(deriv)
res = {}
for mode in ("0", "1"):
    os.environ["QEXXC_I8"] = mode
    rho = ctx.eval_rho(wl.dm, ncomp=wl.ncomp, hermi=1)
    out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
    bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        rho = ctx.eval_rho(wl.dm, ncomp=wl.ncomp, hermi=1)
    e1.record()
    torch.cuda.synchronize()
    e2, e3, e4 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e2.record()
    for _ in range(3):
        out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
    e3.record()
    for _ in range(3):
        bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
    e4.record()
    torch.cuda.synchronize()
    res[mode] = (rho.cpu().numpy(), out.cpu().numpy(), bar.cpu().numpy(), e0.elapsed_time(e1) / 3, e2.elapsed_time(e3) / 3, e3.elapsed_time(e4) / 3)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
r0, o0, b0, t0, f0, v0 = res["0"]
r1, o1, b1, t1, f1, v1 = res["1"]
print(json.dumps({"cfg": cfg, "G": G, "rho_rel": rel(r1, r0), "vmat_rel": rel(o1[0][: N * N], o0[0][: N * N]),
                  "excsum_abs": abs(o1[0][N * N] - o0[0][N * N]), "dm_bar_rel": rel(b1[: N * N], b0[: N * N]),
                  "theta_bar_rel": rel(b1[N * N:], b0[N * N:]), "eval_rho_ms_dmma": round(t0, 3), "eval_rho_ms_i8": round(t1, 3),
                  "fwd_ms_dmma": round(f0, 3), "fwd_ms_i8": round(f1, 3), "vjp_ms_dmma": round(v0, 3), "vjp_ms_i8": round(v1, 3),
                  "finite": bool(np.isfinite(r1).all())}))
