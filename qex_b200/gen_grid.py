"""Atom-centred integration grids (``Grids.coords`` [G,3] Bohr, ``Grids.weights`` [G]).

The reference takes its grid from pyscf: ``grids = pyscf.dft.gen_grid.Grids(mol); grids.level = 0;
grids.becke_scheme = pyscf.dft.gen_grid.stratmann; grids.build()`` (qedft/train/td/trainer_legacy_no_jit.py:
248-251, :316-317; data_io/td/dataset_generation.py:139-142); the hot path only reads ``.coords`` and
``.weights``.  ``Grids(mol)`` here keeps that interface (``level``, ``becke_scheme``, ``prune``, ``atom_grid``,
``build()``) and builds the same grid -- Treutler-Ahlrichs radial nodes, NWChem-pruned Lebedev shells,
Becke / Stratmann fuzzy cells with Treutler's size adjustment.  With the settings above it gives the 1240 points
the reference's notebook logs for H2 and, through the CUDA path, its LDA-RKS energy (tests/test_zz_pyscf_pin.py).

The radial / angular tables are a few hundred numbers per element and are built on the host (as pyscf does); the
O(G natm^2) partition runs either on the host in NumPy (``build()``: set-up code, the way pyscf runs it in C) or
on the GPU (``build(device=0)`` -> ``qexxc_becke_partition``, csrc/grid.cu; raises without the library or a GPU).
A real pyscf ``Grids`` object can be passed to ``qex_b200.numint`` unchanged.

``Grids(mol, n_rad=.., n_theta=.., n_phi=..)`` is the older synthetic product grid (Gauss-Chebyshev radial x
Gauss-Legendre/uniform angular, plain Becke cells) that the synthetic workloads and fixtures are frozen on.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

BOHR = 0.52917721092
# Bragg-Slater radii in Angstrom for Z = 0..10 (index 0: ghost atom)
BRAGG_RADII = np.array((0.35, 0.35, 1.40, 1.45, 1.05, 0.85, 0.70, 0.65, 0.60, 0.50, 1.50)) / BOHR
# xi of the Treutler-Ahlrichs M4 map (JCP 102, 346 (1995), table 1), Z = 0..10
TREUTLER_XI = np.array((1.0, 0.8, 0.9, 1.8, 1.4, 1.3, 1.1, 0.9, 0.9, 0.9, 0.9))
# radial points / Lebedev order by level (rows) and period (H-He, Li-Ne)
_N_RAD = np.array(((10, 15), (30, 40), (40, 60), (50, 75), (60, 90), (70, 105), (80, 120), (90, 135), (100, 150), (200, 200)))
_ANG_ORDER = np.array(((11, 15), (17, 23), (23, 29), (29, 29), (35, 41), (41, 47), (47, 53), (53, 59), (59, 59), (65, 65)))
_LEB_NGRID = np.array((1, 6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194, 230, 266, 302, 350, 434, 590, 770, 974, 1202, 1454))
_LEB_ORDER = np.array((0, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 35, 41, 47, 53, 59, 65))


_SYMBOLS = ("X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne")


def _charge_of(key) -> int:
    return _SYMBOLS.index(key) if isinstance(key, str) else int(key)


# ---------------------------------------------------------------- switching functions (pyscf names)
def original_becke(g):
    """Becke's three-fold iterated p(x) = (3x - x^3)/2."""
    g = np.asarray(g, dtype=np.float64)
    for _ in range(3):
        g = 1.5 * g - 0.5 * g**3
    return g


def stratmann(g):
    """Stratmann, Scuseria, Frisch, CPL 257, 213 (1996), eq. 14 with a = 0.64."""
    g = np.asarray(g, dtype=np.float64)
    t = g / 0.64
    t2 = t * t
    poly = t * (35.0 + t2 * (-35.0 + t2 * (21.0 - 5.0 * t2))) / 16.0
    return np.where(np.abs(g) >= 0.64, np.sign(g), poly)


# ---------------------------------------------------------------- atomic grids (host)
def treutler_ahlrichs(n, chg=1):
    """M4 radial map -> (r ascending, dr)."""
    k = np.arange(n, 0, -1)  # descending k = ascending r
    theta = k * np.pi / (n + 1)
    x = np.cos(theta)
    scale = TREUTLER_XI[chg] / np.log(2.0)
    lg = np.log((1.0 - x) / 2.0)
    pw = (1.0 + x) ** 0.6
    r = -scale * pw * lg
    drdx = scale * pw * (1.0 / (1.0 - x) - 0.6 * lg / (1.0 + x))
    return r, np.pi / (n + 1) * np.sin(theta) * drdx


def nwchem_prune(nuc, rads, n_ang, radii=BRAGG_RADII):
    """Angular points per radial shell: NWChem's five radial regions around the Bragg radius."""
    rads = np.asarray(rads)
    if n_ang < 50:
        return np.full(rads.shape, n_ang)
    table = _LEB_NGRID[4:]
    if n_ang == 50:
        region = np.array((1, 2, 2, 2, 1))
    else:
        top = int(np.searchsorted(table, n_ang))
        region = np.array((1, 3, top - 1, top, top - 1))
    bounds = {0: (0.25, 0.5, 1.0, 4.5), 1: (0.1667, 0.5, 0.9, 3.5), 2: (0.1, 0.4, 0.8, 2.5)}[0 if nuc <= 2 else (1 if nuc <= 10 else 2)]
    where = np.searchsorted(np.asarray(bounds), rads / (radii[nuc] + 1e-200), side="left")
    return table[region[where]]


def _lebedev(npts):
    from scipy.integrate import lebedev_rule

    x, w = lebedev_rule(int(_LEB_ORDER[int(np.where(_LEB_NGRID == npts)[0][0])]))
    return x.T.copy(), w / (4.0 * np.pi)


def gen_atomic_grids(charges, level=3, prune=nwchem_prune, atom_grid=None):
    """{Z: (coords [n,3] around the origin, vol [n])} for every distinct element."""
    out = {}
    for z in sorted(set(int(c) for c in charges)):
        if atom_grid and z in atom_grid:
            n_rad, n_ang = atom_grid[z]
        else:
            period = 0 if z <= 2 else 1
            n_rad = int(_N_RAD[level, period])
            n_ang = int(_LEB_NGRID[np.where(_LEB_ORDER == _ANG_ORDER[level, period])[0][0]])
        r, dr = treutler_ahlrichs(n_rad, z)
        shell_w = 4.0 * np.pi * r * r * dr
        nang = prune(z, r, n_ang) if callable(prune) else np.full(n_rad, n_ang)
        cs, vs = [], []
        for n in np.unique(nang):
            xyz, wa = _lebedev(n)
            sel = np.nonzero(nang == n)[0]
            cs.append((xyz[:, None, :] * r[sel][None, :, None]).reshape(-1, 3))
            vs.append((wa[:, None] * shell_w[sel][None, :]).reshape(-1))
        out[z] = (np.concatenate(cs), np.concatenate(vs))
    return out


def treutler_atomic_radii_adjust(charges, atomic_radii=BRAGG_RADII):
    """a[i,j] of nu = mu + a (1 - mu^2), from the square roots of the Bragg radii, clipped to |a| <= 1/2."""
    s = np.sqrt(atomic_radii[np.asarray(charges, dtype=int)]) + 1e-200
    chi = s[:, None] / s[None, :]
    return np.clip(0.25 * (1.0 / chi - chi), -0.5, 0.5)


def group_grids_by_boxes(centers, coords, box_size=1.2, margin=4.2):
    """Index array of pyscf's point order (``arg_group_grids``): boxes of about ``box_size`` Bohr over the molecule's
    bounding box grown by ``margin``, one overflow layer per side, boxes visited in lexicographic (x, y, z) order, points
    of a box in generation order."""
    lo = centers.min(axis=0) - margin
    span = centers.max(axis=0) + margin - lo
    nbox = np.rint(span / box_size).astype(np.int64)
    cell = np.floor((coords - lo) / (span / nbox)).astype(np.int64)
    cell = np.minimum(np.maximum(cell, -1), nbox)             # overflow layers -1 and nbox
    key = ((cell[:, 0] + 1) * (nbox[1] + 2) + (cell[:, 1] + 1)) * (nbox[2] + 2) + (cell[:, 2] + 1)
    return np.argsort(key, kind="stable")


# ---------------------------------------------------------------- partition
def _partition_host(coords, owner, vol, centers, adjust, scheme):
    na = centers.shape[0]
    if na == 1:
        return vol.copy()
    chunk = max(32, min(4096, 4_000_000 // (na * na)))  # [chunk, na, na] temporaries
    R = np.linalg.norm(centers[:, None] - centers[None], axis=-1)
    np.fill_diagonal(R, np.inf)
    out = np.empty_like(vol)
    for lo in range(0, coords.shape[0], chunk):
        c = coords[lo : lo + chunk]
        d = np.linalg.norm(c[:, None, :] - centers[None], axis=-1)           # [g, atom]
        mu = (d[:, :, None] - d[:, None, :]) / R[None]                       # [g, i, j]
        if adjust is not None:
            mu = mu + adjust[None] * (1.0 - mu * mu)
        cell = 0.5 * (1.0 - scheme(mu))
        cell[:, np.arange(na), np.arange(na)] = 1.0
        P = cell.prod(axis=2)
        out[lo : lo + chunk] = vol[lo : lo + chunk] * P[np.arange(c.shape[0]), owner[lo : lo + chunk]] / P.sum(axis=1)
    return out


def _partition_cuda(coords, owner, vol, centers, adjust, scheme, device):
    import torch

    from . import _lib

    if scheme is original_becke:
        sid = 0
    elif scheme is stratmann:
        sid = 1
    else:
        raise ValueError("the CUDA partition knows gen_grid.original_becke and gen_grid.stratmann")
    if not torch.cuda.is_available():
        raise RuntimeError("Grids.build(device=...) needs a CUDA device (no CPU fallback behind this call)")
    lib = _lib.load()
    dev = torch.device("cuda", int(device))
    with torch.cuda.device(dev):
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)  # noqa: E731
        c, o, v = t(coords, torch.float64), t(owner, torch.int32), t(vol, torch.float64)
        ac = t(centers, torch.float64)
        adj = None if adjust is None else t(adjust, torch.float64)
        na = centers.shape[0]
        work = torch.empty(na * na, dtype=torch.float64, device=dev)
        w = torch.empty_like(v)
        p = lambda x: C.c_void_p(0 if x is None else x.data_ptr())  # noqa: E731
        _lib.check(lib.qexxc_becke_partition(dev.index, p(c), C.c_long(c.shape[0]), p(o), p(v), p(ac), p(adj), na, sid,
                                             p(work), p(w), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return w.cpu().numpy()


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
    """pyscf-style grid container.  ``Grids(mol)`` + attributes + ``build()``; or ``coords=``/``weights=`` given."""

    def __init__(self, mol=None, n_rad=None, n_theta=8, n_phi=8, coords=None, weights=None):
        self.mol = mol
        self.level = 3                      # pyscf's default; the reference sets 0
        self.becke_scheme = original_becke  # the reference sets stratmann
        self.prune = nwchem_prune
        self.atom_grid = {}                 # {Z: (n_rad, n_ang)} overrides the level tables
        self.radii_adjust = treutler_atomic_radii_adjust
        self.n_rad, self.n_theta, self.n_phi = n_rad, n_theta, n_phi  # n_rad given: synthetic product grid
        self.coords = None if coords is None else np.asarray(coords, dtype=np.float64)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)

    def build(self, mol=None, with_non0tab=False, sort_grids=None, device=None, **kwargs):
        """pyscf's ``Grids.build(mol=None, with_non0tab=False, sort_grids=True)`` plus ``device``:
        ``device=None`` partitions on the host (NumPy), ``device=k`` on GPU k (csrc/grid.cu).  ``sort_grids=True`` gives
        pyscf's box ordering of the points; left unset the points keep generation order (atom by atom), which is what
        the fixtures are frozen on -- every consumer on the path sums over points.  ``with_non0tab`` is accepted and
        ignored (no screening table is built, no zero-weight padding points are appended)."""
        del with_non0tab, kwargs
        if mol is not None:
            self.mol = mol
        if self.n_rad is not None:
            return self._build_product()
        charges = np.asarray(self.mol.atom_charges(), dtype=int)
        centers = np.asarray(self.mol.atom_coords(), dtype=np.float64)
        atom_grid = {_charge_of(k): tuple(v) for k, v in (self.atom_grid or {}).items()}  # pyscf keys by symbol
        tab = gen_atomic_grids(charges, self.level, self.prune, atom_grid)
        coords = np.concatenate([tab[int(z)][0] + centers[ia] for ia, z in enumerate(charges)])
        vol = np.concatenate([tab[int(z)][1] for z in charges])
        owner = np.concatenate([np.full(tab[int(z)][1].shape[0], ia, dtype=np.int32) for ia, z in enumerate(charges)])
        adjust = self.radii_adjust(charges) if callable(self.radii_adjust) else None
        if adjust is not None and not adjust.any():
            adjust = None
        if device is None:
            weights = _partition_host(coords, owner, vol, centers, adjust, self.becke_scheme)
        else:
            weights = _partition_cuda(coords, owner, vol, centers, adjust, self.becke_scheme, device)
        if sort_grids:
            idx = group_grids_by_boxes(centers, coords)
            coords, weights = coords[idx], weights[idx]
        self.coords, self.weights = coords, weights
        return self

    def _build_product(self):
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
