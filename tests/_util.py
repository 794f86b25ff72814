"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import numpy as np


def synth_problem(N, G, C=1, B=1, seed=0, rank=None):
    """ao [B,C,G,N] (N(0,1) with a radial-like envelope), symmetric-ish dm [B,N,N], weights [B,G]."""
    rng = np.random.default_rng(seed)
    env = np.exp(-np.abs(rng.standard_normal((B, 1, G, 1))) * 1.5)
    ao = rng.standard_normal((B, C, G, N)) * env * 0.6
    rank = rank or max(1, N // 4)
    Cm = rng.standard_normal((B, N, rank)) / np.sqrt(N)
    dm = 2.0 * np.einsum("bik,bjk->bij", Cm, Cm)
    dm = dm + 1e-3 * rng.standard_normal((B, N, N))  # not exactly symmetric: exercises hermi=0
    w = np.abs(rng.standard_normal((B, G))) * 10.0 / G
    w[:, :: max(1, G // 7)] = 0.0  # exact zeros, as in grid tails
    return ao, dm, w


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)
