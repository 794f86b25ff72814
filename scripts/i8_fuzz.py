"""Shape sweep of the INT8 contractions against the FP64 DMMA kernels (same context, QEXXC_I8 toggled per call):
eval_rho / eval_rho_vjp for AO counts and grid sizes around every tile boundary, LDA and GGA, hermi 0/1, plus the MO form.
python scripts/i8_fuzz.py -> one JSON line per case, exit code 1 if any max-norm difference exceeds 1e-10."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qex_b200.engine import XCContext
from tests._util import synth_problem

Ns = [1, 7, 63, 64, 65, 127, 128, 129, 200, 255, 256, 257, 511, 513, 1000, 1023, 1025, 1290]
Gs = [1, 127, 128, 129, 1000, 4095, 4096, 4097, 8200, 12289]
rng = np.random.default_rng(0)
worst, bad = 0.0, 0
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
cases = [(N, Gs[(i * 3 + k) % len(Gs)], (1, 4)[(i + k) % 2], (i + k) % 2) for i, N in enumerate(Ns) for k in range(2)]
for N, G, C, hermi in cases:
    ao, dm, w = synth_problem(N, G, C, seed=N * 7 + G)
    if hermi:
        dm = 0.5 * (dm + dm.transpose(0, 2, 1))
    rb = rng.standard_normal((1, C, G))
    nmo = max(1, min(N, 1 + N // 3))
    Cm = rng.standard_normal((N, nmo)) / np.sqrt(N)
    occ = rng.uniform(-1.0, 2.0, nmo)
    res = {}
    for mode in ("0", "1"):
        os.environ["QEXXC_I8"] = mode
        ctx = XCContext(nao=N, ngrids_max=G, ncomp=C)
        ctx.set_grid(None, w).set_ao(ao, C)
        r = ctx.eval_rho(dm, ncomp=C, hermi=hermi).cpu().numpy()
        d = ctx.eval_rho_vjp(rb, ncomp=C, hermi=hermi).cpu().numpy()
        m = ctx.eval_rho_mo(Cm, occ).cpu().numpy()
        res[mode] = (r, d, m)
        ctx.close()
    e = [rel(res["1"][k], res["0"][k]) for k in range(3)]
    ok = all(np.isfinite(res["1"][k]).all() for k in range(3)) and max(e) <= 1e-10
    worst = max(worst, max(e))
    bad += not ok
    print(json.dumps({"N": N, "G": G, "C": C, "hermi": hermi, "rho": e[0], "dm_bar": e[1], "rho_mo": e[2], "ok": bool(ok)}), flush=True)
print(json.dumps({"cases": len(cases), "failed": bad, "worst": worst}))
sys.exit(1 if bad else 0)
