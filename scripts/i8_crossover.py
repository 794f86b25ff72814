"""Where does the INT8 path start to pay?  eval_rho + eval_rho_vjp (rowquad + wsyrk + every slicing pass they need, AO planes
rebuilt each iteration like a new geometry) on both tensor pipes for several AO counts.  python scripts/i8_crossover.py [G]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qex_b200.engine import XCContext

G = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for N in (128, 192, 256, 320, 384, 512, 768, 1000):
    ao = torch.randn(G, N, dtype=torch.float64, device="cuda") * torch.exp(-3 * torch.rand(G, 1, dtype=torch.float64, device="cuda"))
    dm = torch.randn(N, N, dtype=torch.float64, device="cuda")
    rb = torch.randn(1, 1, G, dtype=torch.float64, device="cuda")
    w = torch.rand(G, dtype=torch.float64, device="cuda")
    res = {}
    for mode in ("0", "1"):
        os.environ["QEXXC_I8"] = mode
        ctx = XCContext(nao=N, ngrids_max=G)
        ctx.set_grid(None, w)
        for it in range(2):
            ctx.set_ao(ao, 1); ctx.eval_rho(dm, 1, 1); ctx.eval_rho_vjp(rb, 1, 1)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for it in range(4):
            ctx.set_ao(ao, 1)
        e[1].record()
        for it in range(4):
            ctx.set_ao(ao, 1)  # invalidates the digit planes: the INT8 path re-slices
            ctx.eval_rho(dm, 1, 1); ctx.eval_rho_vjp(rb, 1, 1)
        e[2].record()
        torch.cuda.synchronize()
        res[mode] = (e[1].elapsed_time(e[2]) - e[0].elapsed_time(e[1])) / 4
        ctx.close()
    print(json.dumps({"N": N, "G": G, "dmma_ms": round(res["0"], 3), "int8_ms": round(res["1"], 3), "speedup": round(res["0"] / res["1"], 2)}), flush=True)
    del ao
