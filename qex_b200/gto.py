"""Minimal ``Mole``: just enough of pyscf's molecule object for the XC hot path.

The reference reads ``mol._atm / mol._bas / mol._env`` (libcint tables), ``mol.nao_nr()``,
``mol.nbas`` and ``mol.ao_loc_nr()`` (qedft/train/td/numint_legacy.py:141-144, eval_gto.py:48-70).
pyscf is not installable here, so this class builds the same tables -- including pyscf's
normalisation of the contraction coefficients (``gto_norm`` and ``_nomalize_contracted_ao`` in
``pyscf.gto.mole.make_bas_env``) -- for hand-specified or synthetic basis sets.  A real pyscf
``Mole`` can be passed to ``qex_b200.numint`` unchanged (duck typing on the attributes above).
Host-side set-up code, not on the hot path.
"""
from __future__ import annotations

import math

import numpy as np

BOHR = 0.52917721092  # pyscf.data.nist.BOHR
PTR_ENV_START = 20
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF = 0, 1, 2, 3, 4, 5, 6
CHARGE_OF, PTR_COORD = 0, 1

_Z = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10}

# basis[symbol] = [[l, (exp, c1, c2, ...), ...], ...]  (pyscf's internal format)
BASIS = {
    "sto-3g": {
        "H": [[0, (3.42525091, 0.15432897), (0.62391373, 0.53532814), (0.16885540, 0.44463454)]],
    },
    "6-31g": {
        "H": [
            [0, (18.7311370, 0.03349460), (2.8253937, 0.23472695), (0.6401217, 0.81375733)],
            [0, (0.1612778, 1.0)],
        ],
    },
}


def gaussian_int(n, alpha):
    n1 = (n + 1) * 0.5
    return math.gamma(n1) / (2.0 * np.asarray(alpha, dtype=np.float64) ** n1)


def gto_norm(l, expnt):
    return 1.0 / np.sqrt(gaussian_int(l * 2 + 2, 2.0 * np.asarray(expnt, dtype=np.float64)))


def _normalize_contracted(l, es, cs):
    ee = es[:, None] + es[None, :]
    ee = gaussian_int(l * 2 + 2, ee)
    s1 = 1.0 / np.sqrt(np.einsum("pi,pq,qi->i", cs, ee, cs))
    return cs * s1


def even_tempered_basis(nshell_by_l, alpha0=0.12, beta=2.6, nprim=2):
    """Synthetic basis: for each l, `n` contracted shells of `nprim` primitives, even-tempered exponents."""
    out = []
    for l, n in enumerate(nshell_by_l):
        for k in range(n):
            exps = [alpha0 * (1.0 + 0.35 * l) * beta ** (k + 0.5 * p) for p in range(nprim)]
            coefs = [1.0 / (1.0 + p) for p in range(nprim)]
            out.append([l] + [(e, c) for e, c in zip(exps, coefs)])
    return out


class Mole:
    """atoms: [(symbol or nuclear charge, (x, y, z))]; unit 'Angstrom' or 'Bohr'."""

    def __init__(self, atom, basis="sto-3g", unit="Angstrom"):
        self.atom = [(a[0], tuple(float(v) for v in a[1])) for a in atom]
        self.basis = basis
        self.unit = unit
        self._built = False
        self.build()

    def _basis_for(self, sym):
        if isinstance(self.basis, str):
            return BASIS[self.basis.lower()][sym]
        if isinstance(self.basis, dict):
            return self.basis[sym]
        return self.basis  # a single shell list used for every atom

    def build(self):
        scale = 1.0 / BOHR if self.unit.lower().startswith("a") else 1.0
        env = [0.0] * PTR_ENV_START
        atm, bas = [], []
        for ia, (sym, xyz) in enumerate(self.atom):
            z = _Z.get(sym, 0) if isinstance(sym, str) else int(sym)
            ptr = len(env)
            env.extend([v * scale for v in xyz])
            env.append(0.0)  # zeta slot, as pyscf's make_atm_env
            atm.append([z, ptr, 1, ptr + 3, 0, 0])
        cache = {}
        for ia, (sym, _) in enumerate(self.atom):
            key = sym
            if key not in cache:
                shells = []
                for sh in self._basis_for(sym if isinstance(sym, str) else "X"):
                    l = int(sh[0])
                    prim = np.asarray(sh[1:], dtype=np.float64)
                    es = prim[:, 0]
                    cs = prim[:, 1:]
                    cs = np.einsum("pi,p->pi", cs, gto_norm(l, es))
                    cs = _normalize_contracted(l, es, cs)
                    pe = len(env)
                    env.extend(es.tolist())
                    pc = len(env)
                    env.extend(cs.T.reshape(-1).tolist())  # [nctr][nprim]
                    shells.append((l, es.size, cs.shape[1], pe, pc))
                cache[key] = shells
            for l, nprim, nctr, pe, pc in cache[key]:
                bas.append([ia, l, nprim, nctr, 0, pe, pc, 0])
        self._atm = np.asarray(atm, dtype=np.int32).reshape(-1, 6)
        self._bas = np.asarray(bas, dtype=np.int32).reshape(-1, 8)
        self._env = np.asarray(env, dtype=np.float64)
        self.natm = len(atm)
        self.nbas = len(bas)
        self._built = True
        return self

    def nao_nr(self):
        return int(((self._bas[:, ANG_OF] * 2 + 1) * self._bas[:, NCTR_OF]).sum())

    @property
    def nao(self):
        return self.nao_nr()

    def ao_loc_nr(self):
        dims = (self._bas[:, ANG_OF] * 2 + 1) * self._bas[:, NCTR_OF]
        return np.concatenate([[0], np.cumsum(dims)]).astype(np.int32)

    def atom_coords(self):
        return np.array([self._env[p : p + 3] for p in self._atm[:, PTR_COORD]])

    def atom_charges(self):
        return self._atm[:, CHARGE_OF].copy()

    @property
    def nelectron(self):
        return int(self._atm[:, CHARGE_OF].sum())


def h2(bond_length=0.74, basis="6-31g"):
    """The README's H2 example geometry (qedft README 3D section; bond length in Angstrom)."""
    return Mole([("H", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, bond_length))], basis=basis)


def synthetic_molecule(natoms, nshell_by_l, spacing=2.6, seed=0, jitter=0.25):
    """Atoms on a jittered cubic lattice (Bohr) with an even-tempered s/p/d basis."""
    rng = np.random.default_rng(seed)
    side = int(math.ceil(natoms ** (1.0 / 3.0)))
    pts = [(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:natoms]
    xyz = np.asarray(pts, dtype=np.float64) * spacing + rng.uniform(-jitter, jitter, (natoms, 3))
    atoms = [(6, tuple(r)) for r in xyz]
    return Mole(atoms, basis=even_tempered_basis(nshell_by_l), unit="Bohr")
