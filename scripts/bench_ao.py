"""Time the AO evaluator (K1) alone on a c5/c5gga-shaped problem:  python scripts/bench_ao.py [ngrids] [c5|c5gga|c3]
Prints ms per launch, GB/s of AO rows written and the fraction of the measured HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qex_b200 import workloads
from qex_b200.engine import XCContext

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "c5"
wl = workloads.make(cfg, ngrids=G)
ctx = XCContext(nao=wl.nao, ngrids_max=G, ncomp=wl.ncomp, net=workloads.net_spec(wl))
ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
deriv = 1 if wl.ncomp == 4 else 0
for _ in range(3):
    ctx.eval_ao(deriv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ctx.eval_ao(deriv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
npad = -(-wl.nao // 32) * 32
byts = 8.0 * npad * wl.ncomp * G
peak = 6551.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"cfg": cfg, "G": G, "P": os.environ.get("QEXXC_AO_P"), "threads": os.environ.get("QEXXC_AO_THREADS"),
                  "ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac_hbm": round(byts / ms / 1e6 / peak, 4)}))
