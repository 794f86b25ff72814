"""Seeded synthetic workloads for the BASELINE.json configs.

Host-side set-up only (numpy): molecules as libcint tables, grids, density matrices, network
parameters and cotangents.  Used by bench.py, __graft_entry__.smoke() and the tests so that all
of them see identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import gen_grid, gto


@dataclass
class Workload:
    name: str
    describe: str
    mol: gto.Mole
    coords: np.ndarray   # [G, 3]
    weights: np.ndarray  # [G]
    dm: np.ndarray       # [N, N]
    xctype: str          # "NN" | "GGA" | "NN-AmplitudeEncoding"
    ncomp: int
    net: dict            # kwargs for engine.NetSpec (kind by name)
    theta: np.ndarray
    e_bar: float
    v_bar: np.ndarray    # [N, N]
    extra: dict = field(default_factory=dict)

    @property
    def nao(self):
        return self.dm.shape[-1]

    @property
    def ngrids(self):
        return self.weights.shape[-1]


def _mlp_theta(sizes, seed=0):
    """Glorot-normal W / N(0,1e-2) b, flat [W.ravel(), b] per Dense (stax defaults; seeded numpy)."""
    rng = np.random.default_rng(seed)
    parts = []
    for fi, fo in zip(sizes[:-1], sizes[1:]):
        parts.append((rng.standard_normal((fi, fo)) * np.sqrt(2.0 / (fi + fo))).ravel())
        parts.append(rng.standard_normal(fo) * 1e-2)
    return np.concatenate(parts)


def _mo(N, rank, seed):
    """`rank` doubly occupied orbitals: dm = C occ C^T with occ = 2."""
    rng = np.random.default_rng(seed)
    Cm = rng.standard_normal((N, rank)) / np.sqrt(N)
    return Cm, np.full(rank, 2.0)


def _dm(N, rank, seed):
    Cm, occ = _mo(N, rank, seed)
    return (Cm * occ) @ Cm.T


def _cotangents(N, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((N, N)) / N
    return 1.0, v


def make(config: str = "c5", ngrids: int | None = None, seed: int = 0) -> Workload:
    """config in {"c1", "c2", "c3", "c4", "c5", "c5gga", "c5w512"}; `ngrids` overrides the grid size (tests, CPU samples)."""
    if config == "c5w512":
        # the c5 system with the 3D trainer's DEFAULT network: flax MLP features [512, 512, 1], gelu, output
        # -scale * swish(u), applied per grid point (trainer_legacy_no_jit.py:96-107,136-140); layer-by-layer path
        wl = make("c5", ngrids=ngrids, seed=seed)
        wl.name = "c5w512"
        wl.describe = (f"synthetic large system: {wl.nao} AOs x {wl.ngrids} grid points, LocalMLP rho->512->512->1 gelu, "
                       "-scale*swish output (the 3D trainer's default network), fwd+VJP")
        wl.net = dict(kind="local_mlp", n_features=1, n_hidden=2, width=512, activation="gelu", out_transform=1)
        wl.theta = _mlp_theta([1, 512, 512, 1], seed)
        return wl
    if config == "c5" or config == "c5gga":
        # ~1000 AOs x 1M grid points, LocalMLP (BASELINE.json configs[4])
        mol = gto.synthetic_molecule(50, (4, 2, 2), seed=seed)  # 50 atoms x (4s 2p 2d = 20 AOs) = 1000
        G = ngrids or 1_000_000
        grid = gen_grid.random_grid(mol, G, seed=seed + 1)
        N = mol.nao_nr()
        gga = config == "c5gga"
        sizes = [2 if gga else 1, 64, 64, 64, 1]
        e_bar, v_bar = _cotangents(N, seed + 3)
        return Workload(
            name=config, describe=f"synthetic large system: {N} AOs x {G} grid points, LocalMLP "
            f"{'(rho,sigma)' if gga else 'rho'}->64->64->64->1 tanh, fwd+VJP",
            mol=mol, coords=grid.coords, weights=grid.weights, dm=_dm(N, 150, seed + 2),
            xctype="GGA" if gga else "NN", ncomp=4 if gga else 1,
            net=dict(kind="local_mlp", n_features=2 if gga else 1, n_hidden=3, width=64, activation="tanh"),
            theta=_mlp_theta(sizes, seed), e_bar=e_bar, v_bar=v_bar,
            extra=dict(zip(("mo_coeff", "mo_occ"), _mo(N, 150, seed + 2))))
    if config == "c3":
        # water-size: ~120 AOs, ~50k grid points, LocalMLP GGA features (configs[2])
        mol = gto.synthetic_molecule(3, (5, 5, 4), seed=seed)  # 3 atoms x 40 AOs
        G = ngrids or 50_000
        grid = gen_grid.random_grid(mol, G, seed=seed + 1)
        N = mol.nao_nr()
        e_bar, v_bar = _cotangents(N, seed + 3)
        return Workload(
            name="c3", describe=f"synthetic water-size system: {N} AOs x {G} grid points, LocalMLP "
            "(rho,sigma)->64->64->64->1 GGA-feature XC, fwd+VJP",
            mol=mol, coords=grid.coords, weights=grid.weights, dm=_dm(N, 5, seed + 2), xctype="GGA", ncomp=4,
            net=dict(kind="local_mlp", n_features=2, n_hidden=3, width=64, activation="tanh"),
            theta=_mlp_theta([2, 64, 64, 64, 1], seed), e_bar=e_bar, v_bar=v_bar)
    if config == "c2":
        # H2 / 6-31G with LocalQNN (6 qubits, 2 layers), ~1240 grid points (configs[1])
        mol = gto.h2(0.74, "6-31g")
        grid = gen_grid.Grids(mol, n_rad=31, n_theta=5, n_phi=4).build()  # 2 x 620 = 1240 points
        if ngrids:
            grid = gen_grid.Grids(mol, coords=grid.coords[:ngrids], weights=grid.weights[:ngrids])
        N = mol.nao_nr()
        rng = np.random.default_rng(seed)
        e_bar, v_bar = _cotangents(N, seed + 3)
        c = np.array([0.35, 0.28, 0.35, 0.28])[:, None]
        return Workload(
            name="c2", describe=f"H2/6-31G: {N} AOs x {grid.size} grid points, LocalQNN 6 qubits x 2 layers, fwd+VJP",
            mol=mol, coords=grid.coords, weights=grid.weights, dm=2.0 * c @ c.T, xctype="NN", ncomp=1,
            net=dict(kind="local_qnn", n_features=1, n_hidden=2, width=6, in_scale=1.0),
            theta=rng.uniform(-0.1, 0.1, 36), e_bar=e_bar, v_bar=v_bar)
    if config == "c1":
        # README 3D H2 example: GlobalMLP (G->64->64->64->1, networks.py:132-138), bond lengths 0.74/0.5/1.5,
        # batch 3, "NN-AmplitudeEncoding" branch (configs[0])
        bonds = [0.74, 0.5, 1.5]
        mols = [gto.h2(b, "6-31g") for b in bonds]
        grids = [gen_grid.Grids(m, n_rad=31, n_theta=5, n_phi=4).build() for m in mols]
        N, G = 4, grids[0].size
        rng = np.random.default_rng(seed)
        cs = rng.standard_normal((3, N)) * 0.4
        e_bar, v_bar = _cotangents(N, seed + 3)
        return Workload(
            name="c1", describe=f"README 3D H2 example: GlobalMLP {G}->64->64->64->1, {N} AOs x {G} grid points, "
            "bond lengths 0.74/0.5/1.5, batch 3, NN-AmplitudeEncoding, fwd+VJP",
            mol=mols[0], coords=np.stack([g.coords for g in grids]), weights=np.stack([g.weights for g in grids]),
            dm=np.stack([2.0 * np.outer(c, c) for c in cs]), xctype="NN-AmplitudeEncoding", ncomp=1,
            net=dict(kind="global_mlp", n_hidden=3, width=64, activation="tanh"),
            theta=_mlp_theta([G, 64, 64, 64, 1], seed), e_bar=e_bar, v_bar=v_bar,
            extra=dict(batch=3, envs=np.stack([m._env for m in mols]), mols=mols))
    if config == "c4":
        # batched H2 dissociation curve: 64 bond lengths, LocalMLP, one XC step of every molecule's SCF
        # iteration in ONE batched launch per stage (configs[3]; the SCF driver itself is out of scope)
        nb = ngrids or 64  # `ngrids` doubles as the batch-size override for this config
        bonds = np.linspace(0.4, 3.0, nb)
        mols = [gto.h2(float(b), "6-31g") for b in bonds]
        grids = [gen_grid.Grids(m, n_rad=31, n_theta=5, n_phi=4).build() for m in mols]
        N = 4
        rng = np.random.default_rng(seed)
        cs = rng.standard_normal((nb, N)) * 0.4
        dms = np.stack([2.0 * np.outer(c, c) for c in cs])
        e_bar, v_bar = _cotangents(N, seed + 3)
        wl = Workload(
            name="c4", describe=f"batched H2 dissociation curve: {nb} bond lengths x ({N} AOs x {grids[0].size} grid "
            "points), LocalMLP rho->64->64->64->1, one XC step (fwd+VJP) of all molecules per launch",
            mol=mols[0], coords=np.stack([g.coords for g in grids]), weights=np.stack([g.weights for g in grids]),
            dm=dms, xctype="NN", ncomp=1,
            net=dict(kind="local_mlp", n_features=1, n_hidden=3, width=64, activation="tanh"),
            theta=_mlp_theta([1, 64, 64, 64, 1], seed), e_bar=e_bar, v_bar=v_bar,
            extra=dict(batch=nb, envs=np.stack([m._env for m in mols]), mols=mols))
        return wl
    raise ValueError(f"unknown workload {config!r}")


def net_spec(wl: Workload, precision: str = "f64"):
    from . import _lib
    from .engine import NetSpec

    kinds = {"local_mlp": _lib.NET_LOCAL_MLP, "global_mlp": _lib.NET_GLOBAL_MLP, "local_qnn": _lib.NET_LOCAL_QNN}
    kw = dict(wl.net)
    kw["kind"] = kinds[kw["kind"]]
    return NetSpec(precision=precision, **kw)
