#!/usr/bin/env python
"""Benchmark of the XC grid-integration hot path (AO evaluation + nr_rks forward + its VJP).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...  # the reference math on the host cores

One "step" = one pass of the hot path over the whole synthetic workload (default c5:
1000 AOs x 1,000,000 grid points, LocalMLP XC, float64): K1 AO values on the grid, stage 2 rho,
stage 3 network exc/vrho, stage 4 E_xc / V_xc, then the reverse pass (dm_bar, theta_bar).  For
N > 1 the grid is sharded in contiguous point ranges (strong scaling: the global grid is fixed)
and the packed outputs are all-reduced with NCCL (one collective per direction).

Prints ONE JSON line (rank 0).  Timing: CUDA events around exactly K steps, barrier +
synchronize on both sides, max over ranks; inputs (8 GB AO tensor per pass) exceed L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "XC grid-integration grid-pts/s (fwd+VJP)"
UNIT = "grid-pts/s"


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c5gga", "c3", "c2", "c4", "c1", "n2jk", "c4scf", "c1train"],
                    help="c5 = the BASELINE headline; n2jk / c4scf = the widening rows (J/K roofline, batched SCF loop)")
    ap.add_argument("--ngrids", type=int, default=None, help="override the grid size (debugging)")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"], help="network precision")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the whole step as one CUDA graph (auto: single GPU and <= 200k grid points, where the "
                         "step is launch-latency bound)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample duration")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle restatement of the reference math on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_step_time(wl, sample, repeats=1):
    """Seconds for one oracle step (AO + fwd + VJP) on the first `sample` grid points."""
    from oracle import step_ref

    m = wl.mol
    best = float("inf")
    if "batch" in wl.extra:  # c4: `sample` counts grid points over whole molecules
        nmol = max(1, min(wl.extra["batch"], sample // wl.ngrids))
        for _ in range(repeats):
            t0 = time.perf_counter()
            for b in range(nmol):
                mb = wl.extra["mols"][b]
                step_ref.xc_step(mb._atm, mb._bas, mb._env, wl.coords[b], wl.weights[b], wl.dm[b], wl.net, wl.theta,
                                 wl.xctype, wl.e_bar, wl.v_bar)
            best = min(best, time.perf_counter() - t0)
        return best
    c, w = wl.coords[:sample], wl.weights[:sample]
    for _ in range(repeats):
        t0 = time.perf_counter()
        step_ref.xc_step(m._atm, m._bas, m._env, c, w, wl.dm, wl.net, wl.theta, wl.xctype, wl.e_bar, wl.v_bar)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(wl, target_s):
    """Bounded sample of the same workload (about `target_s` seconds of CPU work)."""
    G = wl.ngrids * (wl.extra.get("batch") or 1)
    s0 = min(G, 2048)
    t0 = cpu_step_time(wl, s0)  # calibration (also warms BLAS threads)
    sample = int(min(G, max(s0, s0 * target_s / max(t0, 1e-6) / 2)))
    sample = max(256, (sample // 256) * 256) if G >= 256 else G
    t = cpu_step_time(wl, sample)
    return {"value": sample / t, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
            "sample": f"first {sample} of {G} grid points of the same workload, 1 step (AO eval + nr_rks fwd + VJP), "
                      f"NumPy/BLAS float64 oracle restatement of the reference math (real JAX/pyscfad not installable); "
                      f"{t:.2f} s", "host_cpus": os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's own math on the host cores (the oracle port; the
    reference itself -- JAX + pyscfad -- cannot be installed in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qex_b200 import workloads

    wl = workloads.make(args.config, ngrids=args.ngrids)
    G = wl.ngrids * (wl.extra.get("batch") or 1)
    s0 = min(G, 2048)
    t0 = cpu_step_time(wl, s0)
    budget = 120.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    sample = int(min(G, max(s0, s0 * min(budget, args.cpu_seconds) / max(t0, 1e-6) / 2)))
    sample = max(256, (sample // 256) * 256) if G >= 256 else G
    for _ in range(args.warmup):
        cpu_step_time(wl, sample)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_step_time(wl, sample)
    dt = (time.perf_counter() - t) / max(1, args.steps)
    val = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.describe, "nao": wl.nao, "ngrids": G, "sample_ngrids": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
                         "sample": f"each step = first {sample} of {G} grid points (AO eval + nr_rks fwd + VJP), "
                                   "NumPy/BLAS float64 oracle port; the reference's JAX/pyscfad stack is not installable here",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def measure_dgemm_peak(torch, seconds=1.5):
    """cuBLAS DGEMM TFLOP/s: burst (best of 5) and sustained (back to back for `seconds`).
    MEASURED_PEAKS.json has no FP64 figure, so the FP64 denominator is measured in-run."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    fl = 2.0 * n**3
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, int(seconds / (fl / (best * 1e12))))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    e1.synchronize()
    sustained = fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
    return best, sustained


def run_ours(args):
    import torch
    import torch.distributed as dist

    from qex_b200 import workloads
    from qex_b200.engine import XCContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its banner on STDOUT, next to the one JSON line the driver reads
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workloads.make(args.config, ngrids=args.ngrids)
    N, G = wl.nao, wl.ngrids
    batch = wl.extra.get("batch")
    net = workloads.net_spec(wl, args.precision)
    deriv = 1 if wl.ncomp == 4 else 0
    if batch:
        # c4: molecules are sharded round-robin (replicas); only theta_bar crosses ranks
        from qex_b200.dist import shard_batch

        ids = shard_batch(batch, rank, world)
        Bl, Gl, lo, hi = len(ids), G, 0, G
        ctx = XCContext(nao=N, ngrids_max=G, ncomp=wl.ncomp, nbatch=Bl, net=net, device=local)
        ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.extra["envs"][ids])
        npts_total = batch * G
    else:
        # contiguous grid shard of this rank (fixed map rank -> point range)
        per = (G + world - 1) // world
        lo, hi = min(G, rank * per), min(G, (rank + 1) * per)
        Bl, Gl = 1, hi - lo
        ctx = XCContext(nao=N, ngrids_max=max(Gl, 1), ncomp=wl.ncomp, nbatch=1, net=net, device=local)
        ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env)
        npts_total = G

    # pinned host copies of every per-call input, and device-resident copies
    def pin(x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).pin_memory()

    if batch:
        h_coords, h_w, h_dm = pin(wl.coords[ids]), pin(wl.weights[ids]), pin(wl.dm[ids])
        h_eb, h_vb = pin(np.full(Bl, wl.e_bar)), pin(np.broadcast_to(wl.v_bar, (Bl, N, N)))
    else:
        h_coords, h_w, h_dm = pin(wl.coords[lo:hi]), pin(wl.weights[lo:hi]), pin(wl.dm)
        h_eb, h_vb = pin(np.array([wl.e_bar])), pin(wl.v_bar)
    h_th = pin(wl.theta)
    d_coords, d_w, d_dm, d_th, d_eb, d_vb = (t.cuda() for t in (h_coords, h_w, h_dm, h_th, h_eb, h_vb))
    out = ctx.empty(Bl, N * N + 2)
    bar = ctx.empty(Bl * N * N + wl.theta.size)
    resid = ctx.empty(ctx.resid_doubles)
    h_out, h_bar = torch.empty_like(out, device="cpu").pin_memory(), torch.empty_like(bar, device="cpu").pin_memory()

    def step():
        ctx.set_grid(d_coords, d_w)
        ctx.eval_ao(deriv)
        ctx.nr_rks_fwd(d_dm, d_th, wl.xctype, 0, out=out, resid=resid)
        if world > 1 and not batch:
            dist.all_reduce(out)
        ctx.nr_rks_vjp(d_th, resid, d_eb, d_vb, wl.xctype, 0, out=bar)
        if world > 1:
            dist.all_reduce(bar[Bl * N * N:] if batch else bar)

    # Launch-latency-bound sizes (H2, water): capture the ~16 launches of a step into one CUDA graph.
    use_graph = args.graph == "on" or (args.graph == "auto" and world == 1 and npts_total <= 200_000)
    graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()  # warm-up: schedules uploaded, attributes set, nothing left to allocate
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
    run_step = graph.replay if graph is not None else step

    def step_e2e():
        # host buffers in, host buffers out: what a caller of the drop-in nr_rks pays per call
        d_coords.copy_(h_coords, non_blocking=True)
        d_w.copy_(h_w, non_blocking=True)
        d_dm.copy_(h_dm, non_blocking=True)
        d_th.copy_(h_th, non_blocking=True)
        d_eb.copy_(h_eb, non_blocking=True)
        d_vb.copy_(h_vb, non_blocking=True)
        run_step()
        h_out.copy_(out, non_blocking=True)
        h_bar.copy_(bar, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    peak_burst = peak_sus = None
    if rank == 0:
        peak_burst, peak_sus = measure_dgemm_peak(torch)
    for _ in range(max(3, args.warmup)):
        run_step()
    barrier()

    # ---- timed region: value ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if graph is None:
        ctx.profile_enable(True)
        l0 = ctx.launch_count
        ms_total = timed(step, args.steps)
        launches = ctx.launch_count - l0
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        prof_total = ms_total
    else:
        ms_total = timed(run_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if graph is not None:
        # per-kernel events cannot be recorded inside a replayed graph: take the kernel table (and the
        # launch count) from a separate un-graphed pass of the same K steps
        ctx.profile_enable(True)
        l0 = ctx.launch_count
        prof_total = timed(step, args.steps)
        launches = ctx.launch_count - l0
        prof = ctx.profile_read()
        ctx.profile_enable(False)
    ms_step = ms_total / args.steps
    value = npts_total / (ms_step * 1e-3)

    # ---- extra: the same step with the MO form of rho (dm tagged with mo_coeff/mo_occ takes pyscf's
    # eval_rho2 branch, numint_legacy.py:527-545).  Reported separately; the headline stays dense-dm. ----
    mo_ms = None
    if "mo_coeff" in wl.extra and wl.xctype != "GGA":
        d_C, d_occ = ctx.dev(wl.extra["mo_coeff"]), ctx.dev(wl.extra["mo_occ"])

        def step_mo():
            ctx.set_grid(d_coords, d_w)
            ctx.eval_ao(deriv)
            ctx.nr_rks_fwd_mo(d_C, d_occ, d_th, wl.xctype, out=out, resid=resid)
            if world > 1:
                dist.all_reduce(out)
            ctx.nr_rks_vjp(d_th, resid, d_eb, d_vb, wl.xctype, 0, out=bar)
            if world > 1:
                dist.all_reduce(bar)

        for _ in range(2):
            step_mo()
        mo_ms = timed(step_mo, args.steps) / args.steps

    # ---- e2e: host buffers, copies inside the timed region ----
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    h2d = sum(t.numel() * 8 for t in (h_coords, h_w, h_dm, h_th, h_eb, h_vb))
    d2h = (h_out.numel() + h_bar.numel()) * 8

    if rank == 0:
        # dominant kernel: the FP64 DMMA contractions (2 rowquad + 2 wsyrk launches per step,
        # 2*G*N^2 algorithmic FLOP each -> 8*N^2 FLOP per grid point per step)
        fl = 2.0 * Gl * Bl * N * N
        # executed DMMA FLOPs per launch (zero-padded tiles, symmetric operands skipped):
        # (counted by the library with the same predicates the kernels use)
        tri = wl.ncomp == 1
        ex = {"rowquad": ctx.contraction_flops(0, tri), "wsyrk": ctx.contraction_flops(1, tri)}
        kern = {}
        for name in ("rowquad", "wsyrk", "xc_fwd", "xc_vjp", "eval_ao"):
            ms, n = prof[name]
            kern[name] = {"launches": n, "avg_ms": (ms / n) if n else None,
                          "share_of_step": ms / prof_total if prof_total else None}
            if name in ex and n:
                kern[name]["algorithmic_tflops"] = fl / (ms / n * 1e-3) / 1e12
                kern[name]["executed_tflops"] = ex[name] / (ms / n * 1e-3) / 1e12
                kern[name]["executed_frac_of_peak"] = kern[name]["executed_tflops"] / peak_sus if peak_sus else None
        dom = max(("rowquad", "wsyrk"), key=lambda k: prof[k][0])
        avg_ms = kern[dom]["avg_ms"] or float("nan")
        achieved = fl / (avg_ms * 1e-3) / 1e12
        # DRAM bytes per launch from the committed `ncu --set full` capture (only valid for the captured shape)
        traffic = None
        if wl.name == "c5" and world == 1 and G == 1_000_000:
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["c5"].get(dom + "_kernel")
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor", "kernel": dom + "_kernel (FP64 DMMA.8x8x4)", "achieved": achieved, "peak": peak_sus,
            "unit": "TFLOP/s", "frac": achieved / peak_sus if peak_sus else None, "traffic": traffic,
            "traffic_note": "dram read+write bytes per launch (ncu capture in profiles/r01/contract_ncu_raw.csv); the AO "
                            "tensor is 8.2 GB: a CTA re-reads its 1 MB row tile once per column tile and 148 MB of "
                            "concurrent tiles exceed L2, but the kernel is DMMA-bound (DRAM < 10% busy)",
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run, sustained (MEASURED_PEAKS.json has no FP64 figure); "
                           f"burst {peak_burst:.1f} TFLOP/s",
            "algorithmic_flop_per_launch": fl, "executed_flop_per_launch": ex[dom],
            "executed_tflops": kern[dom]["executed_tflops"], "executed_frac": kern[dom]["executed_frac_of_peak"],
            "note": "achieved/frac use SURVEY 8d's ALGORITHMIC 2*G*N^2 FLOP per launch; the operands are symmetric "
                    "(S = sym(dm), V_xc = ao^T diag ao), so the kernels execute only the upper-triangular tiles "
                    "(executed_* fields): frac > 1 is the symmetry saving, executed_frac is the DMMA-pipe efficiency",
            "kernels": kern,
        }
        cpu = None if args.no_cpu_baseline or world > 1 else cpu_baseline(wl, args.cpu_seconds)
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl.describe, "nao": N, "ngrids": G, "ngrids_per_gpu": Gl, "ncomp": wl.ncomp,
                       "batch": batch, "grid_points_per_step": npts_total,
                       "network_precision": args.precision, "parallelism": (f"molecules sharded round-robin x{world} (replicas), NCCL all-reduce of theta_bar only" if batch else
                                       f"grid-sharded x{world}, NCCL all-reduce of packed V_xc|E_xc|nelec and dm_bar|theta_bar"),
                       "l2": "inputs larger than L2 (AO tensor %.1f GB per pass)" % (Gl * ctx_npad(N) * 8 * wl.ncomp / 1e9),
                       "step": "set_grid + eval_ao (K1) + nr_rks fwd + nr_rks VJP",
                       "launch": "one CUDA graph replay per step (kernel table from an un-graphed pass)" if graph is not None
                       else "stream launches"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": npts_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "mo_path": None if mo_ms is None else {
                "value": npts_total / (mo_ms * 1e-3), "unit": UNIT, "ms_per_step": mo_ms,
                "note": "same step with stage 2 in its MO form rho = sum_k occ_k (ao C_k)^2 (150 occupied orbitals), the "
                        "branch the reference takes when dm carries mo_coeff/mo_occ (numint_legacy.py:527-545); not the headline"},
            "gpu_launches": int(launches), "clocks": clocks, "hbm_peak_gbs": hbm,
            "workspace_gb": ctx.workspace_bytes / 1e9,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ctx_npad(N):  # AO row pitch: a multiple of 32 columns (zeros in the pad)
    return ((N + 31) // 32) * 32


def run_widening(args):
    """`--config n2jk` / `--config c4scf`: the measurements of scripts/bench_jk.py and scripts/bench_scf_c4.py
    with their CPU baselines; bench.py is the one place that may time code under oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    if args.config == "n2jk":
        import bench_jk
        from oracle import jk_ref

        Nc = 64
        e_c = np.random.default_rng(0).standard_normal((Nc,) * 4)
        d_c = np.random.default_rng(1).standard_normal((Nc, Nc))
        t0 = time.perf_counter()
        jk_ref.dot_eri_dm(e_c, d_c)
        t_cpu = time.perf_counter() - t0
        cpu = {"value": 8.0 * Nc**4 / t_cpu / 1e9, "unit": "GB/s", "cores": _cpu_threads(), "kind": "port",
               "sample": f"numpy einsum oracle (J and K) on a [{Nc}]^4 tensor"}
        line = bench_jk.measure(bench_jk.parse([]), cpu)
    elif args.config == "c1train":
        import bench_train
        from oracle import mlp_ref, train_ref
        from qex_b200 import gto

        def cpu_train(bonds, cycles, is_global):
            data = train_ref.make_dataset([gto.h2(b, "6-31g") for b in bonds], level=0)
            G = data[0][1].shape[0]
            spec = mlp_ref.MLPSpec([G if is_global else 1, 64, 64, 64, 1], "tanh")
            theta = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
            t0 = time.perf_counter()
            train_ref.batch_loss(theta, spec, data, 1.0, 1.0, is_global=is_global, max_cycle=cycles)
            t = time.perf_counter() - t0
            return {"value": 1.0 / t, "unit": "it/s", "cores": _cpu_threads(), "kind": "port",
                    "sample": "ONE forward evaluation of the batch loss by the numpy restatement (no gradient: the "
                              "reference's step adds a reverse pass, so this over-states its rate)"}

        line = bench_train.measure(bench_train.parse([]), cpu_train)
    else:
        import bench_scf_c4
        from oracle import gto_ref, mlp_ref, scf_ref

        def cpu_fn(mols, grids, I, theta, cycles, nc=4):
            spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
            t0 = time.perf_counter()
            for b in range(nc):
                ao = gto_ref.eval_ao(mols[b]._atm, mols[b]._bas, mols[b]._env, grids[b].coords, 0)
                d0 = scf_ref.core_guess(I[b]["h1e"], I[b]["s1e"], 2)
                scf_ref.scf_loop(d0, I[b]["eri"], ao, grids[b].weights, I[b]["s1e"], I[b]["h1e"], I[b]["enuc"], 2,
                                 lambda rho: mlp_ref.exc_and_vrho_local(spec, theta, rho), max_cycle=cycles)
            s_per_mol = (time.perf_counter() - t0) / nc
            return {"value": grids[0].size * (cycles + 1) / s_per_mol, "unit": "grid-pts/s", "kind": "port",
                    "cores": _cpu_threads(), "s_per_molecule": s_per_mol,
                    "sample": f"numpy oracle scf_loop on {nc} of the {len(mols)} molecules"}

        line = bench_scf_c4.measure(bench_scf_c4.parse([]), cpu_fn)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = _args()
    if a.config in ("n2jk", "c4scf", "c1train"):
        run_widening(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
