"""BASELINE config c4 end to end: the batched H2 dissociation curve -- 64 bond lengths,
LocalMLP functional, a 15-cycle KS-SCF of every molecule -- as ONE batched device-resident loop
(qex_b200.scf.scf_loop_batched: batched XC kernels, batched J kernel, batched eigensolver / DIIS).

    python scripts/bench_scf_c4.py [--nmol 64] [--cycles 15] [--steps 5] [--warmup 2]

Prints ONE JSON line.  Inputs (integrals, grids, core-Hamiltonian guess) are generated on the host once;
the timed region is the SCF loop itself with everything resident on the device, CUDA events on the
launching stream.  `python bench.py --config c4scf` runs the same measurement and adds the CPU baseline
(the numpy oracle loop on a sample of the molecules).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qex_b200 import _lib, gen_grid, gto, ints, scf, workloads  # noqa: E402
from qex_b200.engine import NetSpec, XCContext  # noqa: E402


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmol", type=int, default=64)
    ap.add_argument("--cycles", type=int, default=15)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    return ap.parse_args(argv)


def measure(args, cpu_baseline_fn=None):
    """-> the JSON line as a dict.  `cpu_baseline_fn(mols, grids, ints, theta, cycles)` is supplied by
    `bench.py --config c4scf` (the only place allowed to time the oracle loop)."""
    B = args.nmol
    bonds = np.linspace(0.4, 3.0, B)
    mols = [gto.h2(float(b), "6-31g") for b in bonds]
    grids = [gen_grid.Grids(m, n_rad=31, n_theta=5, n_phi=4).build() for m in mols]
    I = [ints.integrals(m._atm, m._bas, m._env) for m in mols]
    G, N = grids[0].size, 4
    theta = workloads._mlp_theta([1, 64, 64, 64, 1], 0)
    net = NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64)
    xc = XCContext(nao=N, ngrids_max=G, ncomp=1, nbatch=B, net=net)
    xc.set_grid(np.stack([g.coords for g in grids]), np.stack([g.weights for g in grids]))
    xc.set_basis(mols[0]._atm, mols[0]._bas, np.stack([m._env for m in mols])).eval_ao(0)
    st = lambda k: torch.as_tensor(np.stack([x[k] for x in I])).cuda()  # noqa: E731
    eri, s1e, h1e = st("eri"), st("s1e"), st("h1e")
    enuc = torch.as_tensor(np.array([x["enuc"] for x in I])).cuda()
    th = torch.as_tensor(theta).cuda()
    w, c = scf.generalized_eigh_batched(h1e, s1e)
    dm0 = scf.make_rdm1(c, scf.get_occ_batched(2, w))

    def run():
        with torch.no_grad():
            return scf.scf_loop_batched(xc, th, dm0, eri, s1e, h1e, enuc, 2, max_cycle=args.cycles)

    for _ in range(max(2, args.warmup)):
        e, dm, hist = run()
    torch.cuda.synchronize()
    n0 = int(xc.lib.qexxc_launch_count(xc._h)) + int(xc.lib.qexxc_jk_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e, dm, hist = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # the same loop as ONE CUDA graph (possible because neither the eigensolver kernel nor the DIIS solve
    # synchronises with the host)
    ms_graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            eg_, dmg_, histg_ = run()
        gr.replay()
        torch.cuda.synchronize()
        assert torch.equal(histg_, hist)
        e0.record()
        for _ in range(args.steps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / args.steps
    except Exception as ex:  # capture is an optimisation, not a requirement
        print("graph capture failed:", repr(ex)[:300], file=sys.stderr)
    launches = (int(xc.lib.qexxc_launch_count(xc._h)) + int(xc.lib.qexxc_jk_launch_count()) - n0) // args.steps
    # the same loop differentiated w.r.t. theta (what a training step of the reference does)
    thg = th.clone().requires_grad_(True)
    t0 = time.perf_counter()
    eg, _, _ = scf.scf_loop_batched(xc, thg, dm0, eri, s1e, h1e, enuc, 2, max_cycle=args.cycles)
    (g,) = torch.autograd.grad(eg.sum(), thg)
    torch.cuda.synchronize()
    ms_grad = (time.perf_counter() - t0) * 1e3
    cpu = cpu_baseline_fn(mols, grids, I, theta, args.cycles) if cpu_baseline_fn else None
    pts = B * G * (args.cycles + 1)
    line = {
        "metric": "batched KS-SCF (c4): XC grid-point evaluations per second through the whole SCF loop", "unit": "grid-pts/s",
        "value": pts / (ms * 1e-3), "ms_per_scf_batch": ms, "ms_per_cycle": ms / (args.cycles + 1),
        "ms_per_scf_batch_cuda_graph": ms_graph,
        "ms_loop_plus_theta_gradient": ms_grad, "gpu_launches_per_batch": int(launches),
        "config": {"workload": f"c4: {B} H2/6-31G geometries (0.4-3.0 A), LocalMLP 1->64->64->64->1, {args.cycles}-cycle "
                               f"KS-SCF with DIIS, {G} grid points x {N} AOs each, one batched loop", "nmol": B, "cycles": args.cycles},
        "dtype": "f64", "final_energy_min_max": [float(e.min()), float(e.max())],
        "cpu_baseline": cpu,
    }
    return line


if __name__ == "__main__":
    print(json.dumps(measure(parse())))
