"""Time the two contraction kernels in isolation on random AO data:
python scripts/kernel_probe.py N G [reps]  -> ms and executed TFLOP/s per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from qex_b200.engine import XCContext

for N in [int(x) for x in sys.argv[1].split(",")]:
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = XCContext(nao=N, ngrids_max=G)
    ctx.set_grid(None, torch.rand(G, dtype=torch.float64, device="cuda"))
    ao = torch.randn(G, N, dtype=torch.float64, device="cuda")
    ctx.set_ao(ao, 1)
    dm = torch.randn(N, N, dtype=torch.float64, device="cuda")
    ctx.eval_rho(dm, 1, 0)  # builds the triangular S operand
    for which, name in ((0, "rowquad"), (1, "wsyrk"), (2, "rowquad")):
        for _ in range(2):
            ctx.debug_run_contraction(which)
        torch.cuda.synchronize()
        ctx.profile_enable(True)
        for _ in range(reps):
            ctx.debug_run_contraction(which)
        pr = ctx.profile_read()[name]
        ctx.profile_enable(False)
        ms = pr[0] / pr[1]
        ex = ctx.contraction_flops(0 if which == 2 else which, which != 2)
        print(f"N={N} G={G} {name}{'(dense)' if which == 2 else ''}: {ms:.3f} ms  executed {ex / ms / 1e9:.2f} TF/s  algorithmic {2.0 * G * N * N / ms / 1e9:.2f} TF/s",
              flush=True)
    ctx.close()
    del ao
    torch.cuda.empty_cache()
