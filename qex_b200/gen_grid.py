"""Minimal atom-centred integration grids (``Grids.coords`` [G,3], ``Grids.weights`` [G]).

The reference takes its grid from ``pyscf.dft.gen_grid.Grids`` (level 0, Stratmann-Becke:
qedft/train/td/trainer_legacy_no_jit.py:248-251, dataset_generation.py:139-142); the hot path
only reads ``.coords`` and ``.weights``.  pyscf is not installable here, so this module builds
grids of the same kind (Gauss-Chebyshev radial x product angular quadrature, Becke fuzzy-cell
partition) for synthetic workloads.  A real pyscf ``Grids`` object can be passed unchanged.
Host-side set-up code, not on the hot path.
"""
from __future__ import annotations

import numpy as np


def _radial_gauss_chebyshev(n, rm=1.0):
    """Becke's mapping of Gauss-Chebyshev (2nd kind) nodes to [0, inf)."""
    i = np.arange(1, n + 1)
    x = np.cos(i * np.pi / (n + 1))
    r = rm * (1 + x) / (1 - x)
    w = np.pi / (n + 1) * np.sin(i * np.pi / (n + 1)) ** 2
    w = w * 2 * rm / ((1 - x) ** 2 * np.sqrt(1 - x * x))
    return r, w * r * r


def _angular_product(nth, nph):
    ct, wt = np.polynomial.legendre.leggauss(nth)
    ph = (np.arange(nph) + 0.5) * 2 * np.pi / nph
    st = np.sqrt(1 - ct * ct)
    xyz = np.stack([np.outer(st, np.cos(ph)), np.outer(st, np.sin(ph)), np.outer(ct, np.ones(nph))], -1).reshape(-1, 3)
    w = np.outer(wt, np.full(nph, 2 * np.pi / nph)).reshape(-1)
    return xyz, w


def _becke_weights(coords, centers, ia):
    """Becke fuzzy-cell weight of atom `ia` at `coords` (no atomic-size adjustment)."""
    na = centers.shape[0]
    if na == 1:
        return np.ones(coords.shape[0])
    d = np.linalg.norm(coords[:, None, :] - centers[None, :, :], axis=-1)  # [G, na]
    R = np.linalg.norm(centers[:, None, :] - centers[None, :, :], axis=-1)
    P = np.ones((coords.shape[0], na))
    for i in range(na):
        for j in range(na):
            if i == j:
                continue
            mu = (d[:, i] - d[:, j]) / R[i, j]
            f = mu
            for _ in range(3):
                f = 1.5 * f - 0.5 * f**3
            P[:, i] *= 0.5 * (1 - f)
    return P[:, ia] / P.sum(1)


class Grids:
    """coords/weights container; ``build()`` fills them from ``mol`` when given."""

    def __init__(self, mol=None, n_rad=20, n_theta=8, n_phi=8, coords=None, weights=None):
        self.mol = mol
        self.n_rad, self.n_theta, self.n_phi = n_rad, n_theta, n_phi
        self.coords = None if coords is None else np.asarray(coords, dtype=np.float64)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)

    def build(self):
        centers = self.mol.atom_coords()
        r, wr = _radial_gauss_chebyshev(self.n_rad)
        ang, wa = _angular_product(self.n_theta, self.n_phi)
        cs, ws = [], []
        for ia in range(centers.shape[0]):
            c = centers[ia] + (r[:, None, None] * ang[None, :, :]).reshape(-1, 3)
            w = (wr[:, None] * wa[None, :]).reshape(-1)
            w = w * _becke_weights(c, centers, ia)
            cs.append(c)
            ws.append(w)
        self.coords = np.concatenate(cs)
        self.weights = np.concatenate(ws)
        return self

    @property
    def size(self):
        return 0 if self.weights is None else self.weights.shape[0]


def random_grid(mol, ngrids, seed=1, extent=3.0):
    """Becke-like synthetic grid for large benchmarks: points scattered around the
    atoms with a radial log-normal law, positive weights spanning several decades."""
    rng = np.random.default_rng(seed)
    centers = mol.atom_coords()
    ia = rng.integers(0, centers.shape[0], ngrids)
    rad = np.exp(rng.normal(0.0, 0.9, ngrids)) * 0.8
    rad = np.minimum(rad, extent * 3)
    u = rng.standard_normal((ngrids, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    coords = centers[ia] + rad[:, None] * u
    weights = 4 * np.pi * rad**3 * 0.9 / (ngrids / centers.shape[0]) * np.exp(rng.normal(0, 0.3, ngrids))
    return Grids(mol, coords=coords, weights=np.abs(weights))
