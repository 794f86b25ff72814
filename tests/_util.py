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


ELEM_FLOOR = 1e-6  # element-wise denominators are floored at this fraction of max|ref|
REPORT = []        # (test id, max-norm relative error, element-wise relative error) of every comparison made


def elem_err(a, b, floor_frac=ELEM_FLOOR):
    """Element-wise relative error max_i |a_i - b_i| / max(|b_i|, floor_frac * max|b|) -- the figure
    `north_star` words as "elements within 1e-10 relative"; the floor keeps exact zeros and elements that are
    pure cancellation residue from dividing by ~0."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mx = max(float(np.abs(b).max()) if b.size else 0.0, 1e-300)
    if b.size == 0:
        return 0.0
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor_frac * mx)).max())


def rel_err(a, b):
    """Max-norm relative error max|a - b| / max|b|.  Every call also records the element-wise figure
    (`elem_err`) next to it; tests/conftest.py writes the table to gpurun_out/parity_report.json."""
    import os

    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    r = float(np.abs(a - b).max() / den) if b.size else 0.0
    try:
        REPORT.append((os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], r, elem_err(a, b)))
    except Exception:
        pass
    return r
