"""Roofline measurement of the incore J/K build (csrc/jk.cu).

    python scripts/bench_jk.py [--nao 120] [--steps 20] [--warmup 3]

Prints ONE JSON line: J+K, J-only and reverse-mode times (CUDA events on the launching stream),
algorithmic bytes = 8*nao^4 (the tensor is read exactly once per call) against the HBM peak in
MEASURED_PEAKS.json, and the library GEMV (torch.mv -> cuBLAS) on the same tensor for J alone.
`python bench.py --config n2jk` runs the same measurement and adds the CPU baseline (numpy-einsum oracle).  The tensor (1.66 GB at nao = 120) is
far larger than the 126 MB L2, so every step streams from HBM.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qex_b200 import hf  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nao", type=int, default=120)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    return ap.parse_args(argv)


def measure(args, cpu_baseline=None):
    """-> the JSON line as a dict.  `cpu_baseline` is filled in by `bench.py --config n2jk` (the only place
    allowed to time the oracle); run directly, this script reports the GPU side only."""
    N = args.nao
    g = torch.Generator("cuda").manual_seed(0)
    naux = 64
    L = torch.randn(naux, N * N, dtype=torch.float64, device="cuda", generator=g) / naux**0.5
    eri = (L.T @ L).reshape((N,) * 4).contiguous()
    del L
    dm = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
    a, b = torch.randn_like(dm), torch.randn_like(dm)
    nbytes = 8.0 * N**4
    n0 = hf.jk_launch_count()
    t_jk = timed(lambda: hf._dot_eri_dm_s1(eri, dm, True, True), args.steps, max(3, args.warmup))
    launches = (hf.jk_launch_count() - n0) // (args.steps + max(3, args.warmup))
    t_j = timed(lambda: hf._dot_eri_dm_s1(eri, dm, True, False), args.steps, max(3, args.warmup))
    t_k = timed(lambda: hf._dot_eri_dm_s1(eri, dm, False, True), args.steps, max(3, args.warmup))
    t_vjp = timed(lambda: hf._dot_eri_dm_s1_vjp(eri, N, a, b), args.steps, max(3, args.warmup))
    t_vjp_j = timed(lambda: hf._dot_eri_dm_s1_vjp(eri, N, a, None), args.steps, max(3, args.warmup))
    E2 = eri.reshape(N * N, N * N)
    dmt = dm.T.contiguous().reshape(-1)
    t_mv = timed(lambda: torch.mv(E2.T, dmt), args.steps, max(3, args.warmup))  # J via cuBLAS GEMV (transposed)
    t_copy = timed(lambda: E2[: N * N // 2].clone(), args.steps, max(3, args.warmup))  # read half + write half
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if peak else "B200_PROFILING.md fallback"
    peak = peak or 6500.0
    gbs = lambda ms: nbytes / (ms * 1e-3) / 1e9  # noqa: E731
    line = {
        "metric": "incore J/K build, ERI bytes streamed per second", "unit": "GB/s", "value": gbs(t_jk),
        "ms_per_call": {"jk": t_jk, "j_only": t_j, "k_only": t_k, "vjp_jk": t_vjp, "vjp_j_only": t_vjp_j,
                        "cublas_gemv_j_only": t_mv, "torch_clone_same_bytes": t_copy},
        "config": {"workload": f"N2: dense s1 ERI [{N}]^4 fp64 ({nbytes / 1e9:.2f} GB, > L2) x 1 density matrix, J and K in one pass",
                   "nao": N, "nset": 1},
        "dtype": "f64", "gpu_launches_per_call": int(launches),
        "roofline": {"bound": "hbm", "achieved": gbs(t_jk), "peak": peak, "unit": "GB/s", "frac": gbs(t_jk) / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": nbytes, "traffic": None,
                     "j_only_frac": gbs(t_j) / peak, "vjp_frac": gbs(t_vjp) / peak,
                     "cublas_gemv_frac": gbs(t_mv) / peak},
        "cpu_baseline": cpu_baseline,
    }
    return line


if __name__ == "__main__":
    print(json.dumps(measure(parse())))
