#!/usr/bin/env python
"""Benchmark of the XC grid-integration hot path (AO evaluation + nr_rks forward + its VJP).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...  # the reference math on the host cores

One "step" = one pass of the hot path over the whole synthetic workload (default c5:
1000 AOs x 1,000,000 grid points, LocalMLP XC, float64): K1 AO values on the grid, stage 2 rho,
stage 3 network exc/vrho, stage 4 E_xc / V_xc, then the reverse pass (dm_bar, theta_bar).  For
N > 1 the grid is sharded in contiguous 128-aligned point ranges (`qex_b200.dist.shard_range`; strong
scaling: the global grid is fixed) and the packed outputs are all-reduced over NCCL through the C ABI
(`qexxc_allreduce`), one collective per direction, the forward one overlapped with the VJP.

Prints ONE JSON line (rank 0).  Timing: CUDA events around exactly K steps, barrier +
synchronize on both sides, max over ranks; inputs (8 GB AO tensor per pass) exceed L2.
The line also carries, all timed in the same run: `roofline` (executed DMMA-pipe fraction of the dominant
contraction + the streaming kernels against the HBM peak), `configs` (the other BASELINE configs: c1, c2, c3,
c4, c5gga and c5 with the float32 network, each with value / ms / dominant kernel / CPU baseline), `e2e`
(host buffers in and out) and, for N > 1, `parity_vs_n1` and the collective times.
"""
from __future__ import annotations

import os
import sys

# `--impl reference` times NumPy/BLAS on ALL host cores; torchrun exports OMP_NUM_THREADS=1 to its children,
# which would silently shrink the BLAS pool, so the thread count is pinned before numpy is imported
if "reference" in sys.argv or os.environ.get("QEX_BENCH_CPU_THREADS"):
    _n = os.environ.get("QEX_BENCH_CPU_THREADS") or str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "XC grid-integration grid-pts/s (fwd+VJP)"
UNIT = "grid-pts/s"
SUB_CONFIGS = ("c1", "c2", "c3", "c4", "c5gga", "c5_dmma", "c5gga_dmma", "c5_f32", "c3_f32", "c4_f32", "c5w512")


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
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs)")
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
def _set_cpu_threads():
    """All host cores for BLAS, whatever OMP_NUM_THREADS the launcher exported.  Returns the count in use."""
    want = int(os.environ.get("QEX_BENCH_CPU_THREADS") or os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=want)  # stays in force for the process (not used as a context manager)
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return want


def cpu_step_time(wl, sample, repeats=1):
    """Seconds for one oracle step (AO + fwd + VJP) on the first `sample` grid points."""
    from oracle import step_ref

    m = wl.mol
    best = float("inf")
    if "batch" in wl.extra:  # c1 / c4: `sample` counts grid points over whole molecules
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


def _cpu_sample(wl, target_s, t0, s0):
    G = wl.ngrids * (wl.extra.get("batch") or 1)
    sample = int(min(G, max(s0, s0 * target_s / max(t0, 1e-6) / 2)))
    if "batch" in wl.extra:
        return max(wl.ngrids, (sample // wl.ngrids) * wl.ngrids)
    return max(256, (sample // 256) * 256) if G >= 256 else G


def cpu_baseline(wl, target_s):
    """Bounded sample of the same workload (about `target_s` seconds of CPU work)."""
    cores = _set_cpu_threads()
    G = wl.ngrids * (wl.extra.get("batch") or 1)
    s0 = min(G, 2048) if "batch" not in wl.extra else wl.ngrids
    t0 = cpu_step_time(wl, s0)  # calibration (also warms BLAS threads)
    sample = _cpu_sample(wl, target_s, t0, s0)
    t = cpu_step_time(wl, sample)
    return {"value": sample / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} of {G} grid points of the same workload, 1 step (AO eval + nr_rks fwd + VJP), "
                      f"NumPy/BLAS float64 oracle restatement of the reference math (real JAX/pyscfad not installable); "
                      f"{t:.2f} s", "host_cpus": os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's own math on the host cores (the oracle port; the
    reference itself -- JAX + pyscfad -- cannot be installed in this image).  BLAS runs on every host core even
    under torchrun (which exports OMP_NUM_THREADS=1): pinned at the top of this file and again here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _set_cpu_threads()
    from qex_b200 import workloads

    wl = workloads.make(args.config, ngrids=args.ngrids)
    G = wl.ngrids * (wl.extra.get("batch") or 1)
    s0 = min(G, 2048) if "batch" not in wl.extra else wl.ngrids
    t0 = cpu_step_time(wl, s0)
    budget = 120.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    sample = _cpu_sample(wl, min(budget, args.cpu_seconds), t0, s0)
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
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = first {sample} of {G} grid points (AO eval + nr_rks fwd + VJP), "
                                   "NumPy/BLAS float64 oracle port; the reference's JAX/pyscfad stack is not installable here",
                         "host_cpus": os.cpu_count(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
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


def measure_i8_peak(device):
    """INT8 tensor-core rate (TOP/s, 2 ops per MAC) of this GPU, measured with the library's own tcgen05 issue loop."""
    import ctypes as C
    from qex_b200 import _lib
    v = C.c_double(0)
    lib = _lib.load()
    best = 0.0
    for _ in range(3):
        if lib.qexxc_i8_peak(int(device), C.byref(v)) != 0:
            return None
        best = max(best, v.value / 1e12)
    return best


def ctx_npad(N):  # AO row pitch: a multiple of 32 columns (zeros in the pad)
    return ((N + 31) // 32) * 32


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        return None


def _err_pair(a, b, floor_frac=1e-6):
    """(max-norm relative error, element-wise relative error with the denominator floored at
    floor_frac * max|b|)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    mx = max(float(np.abs(b).max()), 1e-300)
    d = np.abs(a - b)
    return float(d.max() / mx), float((d / np.maximum(np.abs(b), floor_frac * mx)).max())


class Env:
    """Process-level state shared by the measurements of one bench run."""

    def __init__(self, torch, tdist, world, rank, local, comm):
        self.torch, self.tdist, self.world, self.rank, self.local, self.comm = torch, tdist, world, rank, local, comm

    def barrier(self):
        if self.world > 1:
            self.tdist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, k):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.tdist.all_reduce(ms, op=self.tdist.ReduceOp.MAX)
        return float(ms.item())


def measure(env: Env, args, cfg: str, precision: str, steps: int, warmup: int, headline: bool, peak_sus=None,
            cpu_seconds: float = 0.0, peak_i8=None):
    """Times one workload on this process group.  Returns (summary dict for rank 0, extras)."""
    torch = env.torch
    from qex_b200 import workloads
    from qex_b200.dist import ShardedXC, shard_batch, shard_range
    from qex_b200.engine import XCContext

    world, rank, local = env.world, env.rank, env.local
    wl = workloads.make(cfg, ngrids=args.ngrids if headline else None)
    N, G = wl.nao, wl.ngrids
    batch = wl.extra.get("batch")
    net = workloads.net_spec(wl, precision)
    deriv = 1 if wl.ncomp == 4 else 0
    if batch:
        # c1 / c4: molecules are sharded round-robin (replicas); only theta_bar crosses ranks
        ids = shard_batch(batch, rank, world)
        Bl, Gl, lo, hi = len(ids), G, 0, G
        ctx = XCContext(nao=N, ngrids_max=G, ncomp=wl.ncomp, nbatch=max(Bl, 1), net=net, device=local)
        ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.extra["envs"][ids] if Bl else wl.extra["envs"][:1])
        npts_total = batch * G
    else:
        lo, hi = shard_range(G, rank, world)  # fixed 128-aligned rank -> point-range map
        Bl, Gl = 1, hi - lo
        ctx = XCContext(nao=N, ngrids_max=max(Gl, 1), ncomp=wl.ncomp, nbatch=1, net=net, device=local)
        ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env)
        npts_total = G
    sx = ShardedXC(ctx, comm=env.comm)
    nth = wl.theta.size

    def pin(x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).pin_memory()

    if batch:
        sel = ids if Bl else [0]
        h_coords, h_w, h_dm = pin(wl.coords[sel]), pin(wl.weights[sel]), pin(wl.dm[sel])
        h_eb, h_vb = pin(np.full(len(sel), wl.e_bar)), pin(np.broadcast_to(wl.v_bar, (len(sel), N, N)))
    else:
        h_coords, h_w, h_dm = pin(wl.coords[lo:hi]), pin(wl.weights[lo:hi]), pin(wl.dm)
        h_eb, h_vb = pin(np.array([wl.e_bar])), pin(wl.v_bar)
    h_th = pin(wl.theta)
    d_coords, d_w, d_dm, d_th, d_eb, d_vb = (t.cuda() for t in (h_coords, h_w, h_dm, h_th, h_eb, h_vb))
    Bc = ctx.nbatch
    out = ctx.empty(Bc, N * N + 2)
    bar = ctx.empty(Bc * N * N + nth)
    resid = ctx.empty(ctx.resid_doubles)
    h_out, h_bar = torch.empty_like(out, device="cpu").pin_memory(), torch.empty_like(bar, device="cpu").pin_memory()

    def step(record=False):
        ctx.set_grid(d_coords, d_w)
        ctx.eval_ao(deriv)
        if batch:
            ctx.nr_rks_fwd(d_dm, d_th, wl.xctype, 0, out=out, resid=resid)
            ctx.nr_rks_vjp(d_th, resid, d_eb, d_vb, wl.xctype, 0, out=bar)
            if world > 1:
                sx._all_reduce(bar[Bc * N * N:])
        else:
            sx.step(d_dm, d_th, d_eb, d_vb, wl.xctype, out, bar, resid, record=record)

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

    for _ in range(max(3, warmup)):
        run_step()
    env.barrier()

    # ---- timed region: value ----
    sampler = ClockSampler(local) if headline else None
    if rank == 0 and sampler:
        sampler.start()
    if graph is None:
        ctx.profile_enable(True)
        l0, c0 = ctx.launch_count, (env.comm.calls if env.comm else 0)
        ms_total = env.timed(step, steps)
        launches = ctx.launch_count - l0
        ncoll = (env.comm.calls if env.comm else 0) - c0
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        prof_total = ms_total
    else:
        ms_total = env.timed(run_step, steps)
    clocks = sampler.stop() if (rank == 0 and sampler) else None
    if graph is not None:
        # per-kernel events cannot be recorded inside a replayed graph: take the kernel table (and the
        # launch count) from a separate un-graphed pass of the same K steps
        ctx.profile_enable(True)
        l0 = ctx.launch_count
        prof_total = env.timed(step, steps)
        launches = ctx.launch_count - l0
        ncoll = 0
        prof = ctx.profile_read()
        ctx.profile_enable(False)
    ms_step = ms_total / steps
    value = npts_total / (ms_step * 1e-3)

    # collective times of one step (events on the streams the collectives run on), max over ranks
    coll = None
    if world > 1 and not batch:
        # ranks leave the timed region skewed (rank 0 reads the profile tables): line them up and take the second of
        # two recorded steps, otherwise the events time the wait for the slowest rank, not the collective
        for _ in range(2):
            torch.cuda.synchronize()
            env.tdist.barrier()
            step(record=True)
        f_ms, v_ms = sx.collective_ms()
        t = torch.tensor([f_ms, v_ms], dtype=torch.float64, device="cuda")
        env.tdist.all_reduce(t, op=env.tdist.ReduceOp.MAX)
        coll = {"fwd_allreduce_ms": float(t[0]), "vjp_allreduce_ms": float(t[1]),
                "bytes_each": int(out.numel() * 8), "calls_in_timed_region": int(ncoll),
                "note": "NCCL all-reduce of the packed fp64 buffers through qexxc_allreduce; the forward one runs on a "
                        "side stream under the VJP's first contraction, the VJP one closes the step (exposed)"}

    # ---- the same step with the MO form of rho (headline only) ----
    mo_ms = None
    if headline and "mo_coeff" in wl.extra and wl.xctype != "GGA":
        d_C, d_occ = ctx.dev(wl.extra["mo_coeff"]), ctx.dev(wl.extra["mo_occ"])

        def step_mo():
            ctx.set_grid(d_coords, d_w)
            ctx.eval_ao(deriv)
            ctx.nr_rks_fwd_mo(d_C, d_occ, d_th, wl.xctype, out=out, resid=resid)
            sx._all_reduce(out)
            ctx.nr_rks_vjp(d_th, resid, d_eb, d_vb, wl.xctype, 0, out=bar)
            sx._all_reduce(bar)

        for _ in range(2):
            step_mo()
        mo_ms = env.timed(step_mo, steps) / steps

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----
    if batch or graph is not None:
        # (batched molecules, or a launch-latency-bound single-GPU step replayed as one graph)
        def step_e2e():
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

        h2d = sum(t.numel() * 8 for t in (h_coords, h_w, h_dm, h_th, h_eb, h_vb))
        d2h = (h_out.numel() + h_bar.numel()) * 8
        e2e_note = "every rank copies its own inputs in and its outputs out around the step"
        io = None
    else:
        io = sx.make_host_io(1, nth)
        if rank == 0:
            sx.pack_host_inputs(io, wl.dm, wl.theta, wl.e_bar, wl.v_bar)

        def step_e2e():
            sx.step_host(io, h_coords, h_w, d_coords, d_w, wl.xctype, deriv)

        # rank 0 moves the replicated inputs/outputs; every rank moves its own grid shard (bytes of rank 0)
        h2d = (io["n1"] + io["n2"]) * 8 + (h_coords.numel() + h_w.numel()) * 8
        d2h = (out.numel() + bar.numel()) * 8
        e2e_note = ("rank 0 uploads dm|theta|e_bar|v_bar once (NCCL broadcast to the peers over NVLink) and downloads "
                    "out|bar; every rank uploads its own coords/weights shard; cotangent upload and the forward "
                    "result's reduce + download overlap compute")
    for _ in range(2):
        step_e2e()
    ms_e2e = env.timed(step_e2e, steps) / steps
    e2e_same = None
    if io is not None:
        # the pipelined host path must return what the device-resident path returns (bit for bit)
        run_step()
        torch.cuda.synchronize()
        if rank == 0:
            e2e_same = bool(torch.equal(io["h_out"], out.cpu()) and torch.equal(io["h_bar"], bar.cpu()))

    res = None
    if rank == 0:
        fl = 2.0 * Gl * Bc * N * N
        tri = wl.ncomp == 1
        mode = ctx.contraction_mode
        if mode == "int8":  # exact digit split on the INT8 tensor cores: executed INT8 ops against the measured INT8 rate
            ex = {"rowquad": ctx.contraction_i8_ops(0, tri), "wsyrk": ctx.contraction_i8_ops(1, tri)}
            peak_c = peak_i8
        else:
            ex = {"rowquad": ctx.contraction_flops(0, tri), "wsyrk": ctx.contraction_flops(1, tri)}
            peak_c = peak_sus
        kern = {}
        for name in ("rowquad", "wsyrk", "xc_fwd", "xc_vjp", "eval_ao", "stage4", "slice"):
            ms, n = prof[name]
            kern[name] = {"launches": n, "avg_ms": (ms / n) if n else None,
                          "share_of_step": ms / prof_total if prof_total else None}
            if name in ex and n:
                kern[name]["algorithmic_tflops"] = fl / (ms / n * 1e-3) / 1e12
                kern[name]["executed_tflops"] = ex[name] / (ms / n * 1e-3) / 1e12
                kern[name]["executed_frac_of_peak"] = kern[name]["executed_tflops"] / peak_c if peak_c else None
                kern[name]["executed_unit"] = "TOP/s (INT8)" if mode == "int8" else "TFLOP/s (FP64)"
        dom_all = max(kern, key=lambda k: prof[k][0])
        cpu = None
        if cpu_seconds > 0 and world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(wl, cpu_seconds)
        res = {
            "wl": wl, "value": value, "ms_per_step": ms_step, "kern": kern, "ex": ex, "fl": fl, "dominant": dom_all,
            "mode": mode,
            "e2e": {"value": npts_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "note": e2e_note,
                    "bitwise_equal_to_device_path": e2e_same},
            "launches": int(launches), "clocks": clocks, "cpu": cpu, "mo_ms": mo_ms, "coll": coll,
            "graph": graph is not None, "Gl": Gl, "Bl": Bc, "npts_total": npts_total, "batch": batch,
            "workspace_gb": ctx.workspace_bytes / 1e9, "npad": ctx_npad(N),
        }
    extras = {"ctx": ctx, "out": out, "bar": bar, "wl": wl, "step": step}
    return res, extras


def parity_vs_single_rank(env: Env, extras):
    """N > 1: rank 0 recomputes the step on the WHOLE grid by itself and compares it with the sharded,
    all-reduced result every rank holds (|dE_xc|, V_xc / dm_bar / theta_bar: max-norm and element-wise)."""
    torch = env.torch
    from qex_b200 import workloads
    from qex_b200.engine import XCContext

    wl = extras["wl"]
    extras["step"]()
    torch.cuda.synchronize()
    o_n, b_n = extras["out"].cpu().numpy()[0], extras["bar"].cpu().numpy()
    res = None
    if env.rank == 0:
        N, G = wl.nao, wl.ngrids
        c1 = XCContext(nao=N, ngrids_max=G, ncomp=wl.ncomp, nbatch=1, net=extras["ctx"].net, device=env.local)
        c1.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights).eval_ao(1 if wl.ncomp == 4 else 0)
        o1, r1 = c1.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
        b1 = c1.nr_rks_vjp(wl.theta, r1, [wl.e_bar], wl.v_bar, wl.xctype).cpu().numpy()
        o1 = o1.cpu().numpy()[0]
        nn = N * N
        v = _err_pair(o_n[:nn], o1[:nn])
        d = _err_pair(b_n[:nn], b1[:nn])
        t = _err_pair(b_n[nn:], b1[nn:])
        res = {"dE_xc_abs": abs(float(o_n[nn] - o1[nn])), "nelec_abs": abs(float(o_n[nn + 1] - o1[nn + 1])),
               "vxc_maxnorm_rel": v[0], "vxc_elementwise_rel": v[1], "dm_bar_maxnorm_rel": d[0],
               "dm_bar_elementwise_rel": d[1], "theta_bar_maxnorm_rel": t[0], "theta_bar_elementwise_rel": t[1],
               "elementwise_floor": "denominator max(|ref|, 1e-6 max|ref|)",
               "tolerance": "|dE_xc| <= 1e-9 Ha, elements <= 1e-10 of the largest element (max-norm relative); the element-wise figures "
                            "are reported beside it: with the INT8 digit-split contractions (fixed-point operands, default for "
                            "nao >= 256) small V_xc / dm_bar elements carry the same ~1e-12-of-maximum absolute error",
               "ok": bool(abs(o_n[nn] - o1[nn]) <= 1e-9 and max(v[0], d[0], t[0]) <= 1e-10)}
        c1.close()
    env.barrier()
    return res


def run_ours(args):
    import torch
    import torch.distributed as tdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its banner on STDOUT, next to the one JSON line the driver reads
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from qex_b200.dist import Comm

        comm = Comm(rank, world, local)  # the data-path collectives go through the C ABI (qexxc_allreduce)
    env = Env(torch, tdist, world, rank, local, comm)

    peak_burst = peak_sus = peak_i8 = None
    if rank == 0:
        peak_burst, peak_sus = measure_dgemm_peak(torch)
        peak_i8 = measure_i8_peak(local)
    res, extras = measure(env, args, args.config, args.precision, args.steps, args.warmup, True, peak_sus,
                          args.cpu_seconds, peak_i8)
    parity = parity_vs_single_rank(env, extras) if (world > 1 and not extras["wl"].extra.get("batch")) else None
    ctx_ws = res["workspace_gb"] if res else None
    extras["ctx"].close()
    del extras
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, timed in the same run (single GPU; c4 also when sharded by molecule) ----
    subs = {}
    if not args.no_configs and args.ngrids is None and args.config == "c5":
        todo = SUB_CONFIGS if world == 1 else ("c4",)
        for name in todo:
            cfg, prec = (name[:-4], "f32") if name.endswith("_f32") else (name, "f64")
            force_dmma = name.endswith("_dmma")  # the same workload with the contractions on the FP64 tensor pipe
            if force_dmma:
                cfg = name[:-5]
            old_i8 = os.environ.get("QEXXC_I8")
            if force_dmma:
                os.environ["QEXXC_I8"] = "0"
            try:
                r, ex2 = measure(env, args, cfg, prec, args.steps, args.warmup, False, peak_sus,
                                 0.0 if (name.endswith("_f32") or name == "c5w512" or force_dmma) else min(4.0, args.cpu_seconds),
                                 peak_i8)
                ex2["ctx"].close()
                del ex2
                torch.cuda.empty_cache()
                if rank == 0:
                    k = r["kern"][r["dominant"]]
                    subs[name] = {
                        "workload": r["wl"].describe, "network_precision": prec, "value": r["value"], "unit": UNIT,
                        "ms_per_step": r["ms_per_step"], "grid_points_per_step": r["npts_total"],
                        "e2e_value": r["e2e"]["value"], "e2e_ms_per_step": r["e2e"]["ms_per_step"],
                        "dominant_kernel": r["dominant"], "dominant_share_of_step": k["share_of_step"],
                        "dominant_avg_ms": k["avg_ms"],
                        "kernel_shares": {n: v["share_of_step"] for n, v in r["kern"].items() if v["launches"]},
                        "contraction_pipe": r["mode"],
                        "contraction_executed_frac_of_pipe_peak": {n: r["kern"][n].get("executed_frac_of_peak")
                                                                   for n in ("rowquad", "wsyrk")},
                        "gpu_launches": r["launches"], "cpu_baseline": r["cpu"],
                        "launch": "one CUDA graph replay per step" if r["graph"] else "stream launches",
                    }
            except Exception as e:  # a failing side config must not lose the headline line
                if rank == 0:
                    subs[name] = {"error": f"{type(e).__name__}: {e}"}
            finally:
                if force_dmma:
                    if old_i8 is None:
                        os.environ.pop("QEXXC_I8", None)
                    else:
                        os.environ["QEXXC_I8"] = old_i8

    if rank == 0:
        wl, kern, ex, fl = res["wl"], res["kern"], res["ex"], res["fl"]
        N, G = wl.nao, wl.ngrids
        hbm = _hbm_peak()
        dom = max(("rowquad", "wsyrk"), key=lambda k: (kern[k]["avg_ms"] or 0) * (kern[k]["launches"] or 0))
        avg_ms = kern[dom]["avg_ms"] or float("nan")
        achieved_alg = fl / (avg_ms * 1e-3) / 1e12
        traffic = None
        if wl.name == "c5" and world == 1 and G == 1_000_000:
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["c5"].get(dom + "_kernel")
            except Exception:
                traffic = None
        # streaming kernels against the HBM copy peak (algorithmic bytes, SURVEY 8d / DESIGN section 4)
        C_ = wl.ncomp
        stream_bytes = {
            # K1 writes the AO tensor once: 8 B x Npad x C per point (coords in: 24 B)
            "eval_ao": (8.0 * res["npad"] * C_ + 24.0) * res["Gl"] * res["Bl"],
            # stage 4 fwd: rho(C) exc vrho (vgamma) w in, wv(C) out; adjoint: the same in + wvb(C), rho_bar(C) exc_bar
            # vrho_bar (vgamma_bar) out -- summed over the two launches per step and divided by two below
            "stage4": 0.5 * 8.0 * ((C_ + 3 + (C_ == 4) + C_) + (2 * C_ + 3 + (C_ == 4) + C_ + 2 + (C_ == 4))) * res["Gl"] * res["Bl"],
        }
        if res["mode"] == "int8":
            # per step: the row-scaled planes of the geometry (read ao once, write 6 planes; the column maxima come from the same
            # pass) + two weighted operands (read ao, write 6 planes); per launch of the class = that total over its scopes
            npk = ((res["npad"] + 127) // 128) * 128
            per_step = 3 * (8.0 * res["npad"] + 6.0 * npk) * res["Gl"]
            nsl = kern["slice"]["launches"] or 1
            stream_bytes["slice"] = per_step * args.steps / nsl
        streaming = {}
        for name, nbytes in stream_bytes.items():
            k = kern[name]
            if k["launches"]:
                gbs = nbytes / (k["avg_ms"] * 1e-3) / 1e9
                streaming[name] = {"kernel": {"eval_ao": "eval_ao_kernel (K1)", "stage4": "stage4_fwd/vjp kernels",
                                              "slice": "slice_rows (+ column maxima) / slice_cols (INT8 digit planes)"}[name],
                                   "bytes_per_launch": nbytes, "avg_ms": k["avg_ms"], "gbs": gbs,
                                   "frac_of_hbm_peak": gbs / hbm if hbm else None}
        int8 = res["mode"] == "int8"
        peak_c = peak_i8 if int8 else peak_sus
        if int8:
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["c5_int8"].get(dom + "_i8_kernel") \
                    if (wl.name == "c5" and world == 1 and G == 1_000_000) else None
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor",
            "kernel": dom + ("_i8_kernel (tcgen05.mma kind::i8, exact 6 x 6 digit split of the FP64 product)" if int8
                             else "_kernel (FP64 DMMA.8x8x4)"),
            "achieved": kern[dom]["executed_tflops"], "peak": peak_c, "unit": "TOP/s" if int8 else "TFLOP/s",
            "frac": kern[dom]["executed_frac_of_peak"], "traffic": traffic,
            "frac_definition": ("EXECUTED INT8 operations per launch (21 digit products of every scheduled 128x64x128 block, 2 per MAC) "
                                "/ launch time / the INT8 tensor rate measured in this run with the library's own issue loop "
                                "(qexxc_i8_peak; 2 x MEASURED_PEAKS.json's dense bf16 figure would be "
                                f"{2 * (json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops') or float('nan')):.0f}); "
                                "the FP64-equivalent algorithmic figure is under `algorithmic`") if int8 else
                               ("EXECUTED DMMA FLOP per launch / launch time / cuBLAS DGEMM sustained peak measured in this "
                                "run (the pipe fraction); the algorithmic figure is under `algorithmic`"),
            "algorithmic": {"flop_per_launch": fl, "tflops": achieved_alg,
                            "frac_of_fp64_dgemm_peak": achieved_alg / peak_sus if peak_sus else None,
                            "note": "SURVEY 8d's 2*G*N^2 FP64 FLOP per launch over the launch time, i.e. the FP64-equivalent rate; "
                                    "cuBLAS DGEMM sustained peak measured in this run = %.1f TFLOP/s" % (peak_sus or float("nan"))},
            "executed_per_launch": ex[dom],
            "contraction_pipe": res["mode"],
            "traffic_note": "dram read+write bytes per launch of THIS kernel at this shape, from the committed ncu capture "
                            "(profiles/ncu_traffic.json; not measurable inside a timed run); algorithmic bytes = "
                            + (("the A digit planes (6 B per AO value) + the FP64 ao rows of the row-dot epilogue = %.2f GB" % (14.0 * res["npad"] * res["Gl"] / 1e9)
                                if dom == "rowquad" else
                                "the digit planes of both operands read once = %.2f GB" % (12.0 * res["npad"] * res["Gl"] / 1e9)) if int8
                               else "the AO tensor read once = %.2f GB" % (8.0 * res["npad"] * res["Gl"] / 1e9)),
            "peak_source": ("qexxc_i8_peak measured in this run (best of 3)" if int8 else
                            "cuBLAS DGEMM 8192^3 measured in this run, sustained (MEASURED_PEAKS.json has no FP64 figure); "
                            f"burst {peak_burst:.1f} TFLOP/s"),
            "fp64_dgemm_peak_tflops": peak_sus, "int8_peak_tops": peak_i8,
            "streaming": streaming, "hbm_peak_gbs": hbm, "hbm_peak_source": "MEASURED_PEAKS.json (of measured)",
            "kernels": kern,
        }
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.describe, "nao": N, "ngrids": G, "ngrids_per_gpu": res["Gl"], "ncomp": wl.ncomp,
                       "batch": res["batch"], "grid_points_per_step": res["npts_total"],
                       "network_precision": args.precision,
                       "parallelism": (f"molecules sharded round-robin x{world} (replicas), NCCL all-reduce of theta_bar only"
                                       if res["batch"] else
                                       f"grid-sharded x{world} (128-aligned contiguous ranges), NCCL all-reduce of packed "
                                       "V_xc|E_xc|nelec and dm_bar|theta_bar via qexxc_allreduce"),
                       "l2": "inputs larger than L2 (AO tensor %.1f GB per pass)" % (res["Gl"] * res["npad"] * 8 * wl.ncomp / 1e9),
                       "step": "set_grid + eval_ao (K1) + nr_rks fwd + nr_rks VJP",
                       "launch": "one CUDA graph replay per step (kernel table from an un-graphed pass)" if res["graph"]
                       else "stream launches"},
            "roofline": roofline, "cpu_baseline": res["cpu"], "e2e": res["e2e"],
            "mo_path": None if res["mo_ms"] is None else {
                "value": res["npts_total"] / (res["mo_ms"] * 1e-3), "unit": UNIT, "ms_per_step": res["mo_ms"],
                "note": "same step with stage 2 in its MO form rho = sum_k occ_k (ao C_k)^2 (150 occupied orbitals), the "
                        "branch the reference takes when dm carries mo_coeff/mo_occ (numint_legacy.py:527-545); not the headline"},
            "collectives": res["coll"], "parity_vs_n1": parity, "configs": subs,
            "gpu_launches": res["launches"], "clocks": res["clocks"], "hbm_peak_gbs": hbm, "workspace_gb": ctx_ws,
            "parity_note": "oracle = NumPy restatement of the reference (JAX/pyscfad/horqrux not installable here): MLP, QNN "
                           "and l>0 AO arithmetic have no reference-held pin; the jax.ffi shim is written but cannot be "
                           "built or run in this image",
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        tdist.destroy_process_group()


def run_widening(args):
    """`--config n2jk` / `--config c4scf`: the measurements of scripts/bench_jk.py and scripts/bench_scf_c4.py
    with their CPU baselines; bench.py is the one place that may time code under oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    cores = _set_cpu_threads()
    if args.config == "n2jk":
        import bench_jk
        from oracle import jk_ref

        Nc = 64
        e_c = np.random.default_rng(0).standard_normal((Nc,) * 4)
        d_c = np.random.default_rng(1).standard_normal((Nc, Nc))
        t0 = time.perf_counter()
        jk_ref.dot_eri_dm(e_c, d_c)
        t_cpu = time.perf_counter() - t0
        cpu = {"value": 8.0 * Nc**4 / t_cpu / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
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
            return {"value": 1.0 / t, "unit": "it/s", "cores": cores, "kind": "port",
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
                    "cores": cores, "s_per_molecule": s_per_mol,
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
