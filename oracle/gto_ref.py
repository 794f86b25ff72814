"""Oracle: atomic-orbital (AO) values and gradients on a grid.  TEST INFRASTRUCTURE ONLY.

Restates what the reference obtains from pyscf's C evaluator
(``pyscf.gto.eval_gto`` "GTOval_sph_deriv0/1", reached through
``qedft/train/td/eval_gto.py:48-70`` and ``pyscf ... block_loop`` at
``qedft/train/td/numint_legacy.py:292,313``; also ``numint.eval_ao(mol, coords, deriv=0)``
at ``qedft/train/td/trainer_legacy_no_jit.py:273``).

pyscf ^2.9.0 is a third-party dependency absent from /root/reference, so this follows its
published algorithm: a contracted real-spherical Gaussian shell of angular momentum l on
centre A is

    phi_{l m}(r) = S_{l m}(r - A) * sum_p c_p exp(-alpha_p |r - A|^2)

with S_lm the real solid harmonic (orthonormal on the unit sphere; pyscf order: p = x,y,z;
d = xy, yz, z^2, xz, x^2-y^2; f = m=-3..3) and c_p the coefficients as stored in
``mol._env`` (already multiplied by ``gto_norm(l, alpha_p)`` and normalised as a contracted
function by ``pyscf.gto.mole.make_bas_env``).

Layout (same as pyscf): ``ao[g, i]`` for deriv=0 and ``ao[c, g, i]`` (c = value, d/dx, d/dy,
d/dz) for deriv=1; AO index i runs shell by shell, contraction by contraction, m fastest.

parity unpinned: no reference fixture holds AO values; pinned here by orthonormality only.
"""
from __future__ import annotations

import math

import numpy as np

# libcint slots (pyscf/gto/mole.py: ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, PTR_EXP, PTR_COEFF, PTR_COORD)
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF = 0, 1, 2, 3, 4, 5, 6
PTR_COORD = 1

# Real solid harmonics as polynomial tables: list over m of list of (coef, (a, b, c)) meaning
# coef * x^a y^b z^c.  Constants are sqrt((2l+1)/4pi) * standard real solid harmonics.
_S = 0.28209479177387814  # 1/sqrt(4 pi)
_P = 0.4886025119029199  # sqrt(3/(4 pi))
_SOLID = {
    0: [[(_S, (0, 0, 0))]],
    1: [[(_P, (1, 0, 0))], [(_P, (0, 1, 0))], [(_P, (0, 0, 1))]],
    2: [
        [(1.0925484305920792, (1, 1, 0))],
        [(1.0925484305920792, (0, 1, 1))],
        [(0.6307831305050401, (0, 0, 2)), (-0.31539156525252005, (2, 0, 0)), (-0.31539156525252005, (0, 2, 0))],
        [(1.0925484305920792, (1, 0, 1))],
        [(0.5462742152960396, (2, 0, 0)), (-0.5462742152960396, (0, 2, 0))],
    ],
    3: [
        # m=-3: sqrt(35/(32pi)) (3x^2 y - y^3)
        [(1.7701307697799304, (2, 1, 0)), (-0.5900435899266435, (0, 3, 0))],
        # m=-2: sqrt(105/(4pi)) xyz
        [(2.8906114426405543, (1, 1, 1))],
        # m=-1: sqrt(21/(32pi)) y (4z^2 - x^2 - y^2)
        [(1.8281831978578629, (0, 1, 2)), (-0.4570457994644657, (2, 1, 0)), (-0.4570457994644657, (0, 3, 0))],
        # m=0: sqrt(7/(16pi)) z (2z^2 - 3x^2 - 3y^2)
        [(0.7463526651802308, (0, 0, 3)), (-1.1195289977703462, (2, 0, 1)), (-1.1195289977703462, (0, 2, 1))],
        # m=1: sqrt(21/(32pi)) x (4z^2 - x^2 - y^2)
        [(1.8281831978578629, (1, 0, 2)), (-0.4570457994644657, (3, 0, 0)), (-0.4570457994644657, (1, 2, 0))],
        # m=2: sqrt(105/(16pi)) z (x^2 - y^2)
        [(1.4453057213202771, (2, 0, 1)), (-1.4453057213202771, (0, 2, 1))],
        # m=3: sqrt(35/(32pi)) (x^3 - 3 x y^2)
        [(0.5900435899266435, (3, 0, 0)), (-1.7701307697799304, (1, 2, 0))],
    ],
}
LMAX = 3


def gaussian_int(n, alpha):
    """int_0^inf x^n exp(-alpha x^2) dx  (pyscf.gto.mole.gaussian_int)."""
    n1 = (n + 1) * 0.5
    return math.gamma(n1) / (2.0 * np.asarray(alpha, dtype=np.float64) ** n1)


def gto_norm(l, expnt):
    """Radial normalisation of a primitive (pyscf.gto.mole.gto_norm)."""
    return 1.0 / np.sqrt(gaussian_int(l * 2 + 2, 2.0 * np.asarray(expnt, dtype=np.float64)))


def normalize_contracted(l, es, cs):
    """pyscf.gto.mole._nomalize_contracted_ao: cs[nprim, nctr] already times gto_norm."""
    ee = es[:, None] + es[None, :]
    ee = gaussian_int(l * 2 + 2, ee)
    s1 = 1.0 / np.sqrt(np.einsum("pi,pq,qi->i", cs, ee, cs))
    return cs * s1


def nao_nr(bas):
    bas = np.asarray(bas)
    return int(((bas[:, ANG_OF] * 2 + 1) * bas[:, NCTR_OF]).sum())


def ao_loc_nr(bas):
    bas = np.asarray(bas)
    dims = (bas[:, ANG_OF] * 2 + 1) * bas[:, NCTR_OF]
    return np.concatenate([[0], np.cumsum(dims)]).astype(np.int32)


def _poly_and_grad(terms, x, y, z):
    """value and d/dx, d/dy, d/dz of sum coef x^a y^b z^c (arrays over grid)."""

    def pw(v, n):
        return np.ones_like(v) if n == 0 else v**n

    val = np.zeros_like(x)
    gx = np.zeros_like(x)
    gy = np.zeros_like(x)
    gz = np.zeros_like(x)
    for coef, (a, b, c) in terms:
        val += coef * pw(x, a) * pw(y, b) * pw(z, c)
        if a > 0:
            gx += coef * a * pw(x, a - 1) * pw(y, b) * pw(z, c)
        if b > 0:
            gy += coef * b * pw(x, a) * pw(y, b - 1) * pw(z, c)
        if c > 0:
            gz += coef * c * pw(x, a) * pw(y, b) * pw(z, c - 1)
    return val, gx, gy, gz


def eval_ao(atm, bas, env, coords, deriv=0):
    """AO values (deriv=0 -> [G, N]) or values+gradient (deriv=1 -> [4, G, N]), float64."""
    atm = np.asarray(atm)
    bas = np.asarray(bas)
    env = np.asarray(env, dtype=np.float64)
    coords = np.asarray(coords, dtype=np.float64)
    if deriv not in (0, 1):
        raise NotImplementedError("oracle eval_ao: deriv must be 0 or 1")
    G = coords.shape[0]
    N = nao_nr(bas)
    ncomp = 1 if deriv == 0 else 4
    out = np.zeros((ncomp, G, N))
    i0 = 0
    for ib in range(bas.shape[0]):
        ia, l, nprim, nctr = (int(bas[ib, k]) for k in (ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF))
        if l > LMAX:
            raise NotImplementedError(f"oracle eval_ao: l={l} > {LMAX}")
        pc = int(atm[ia, PTR_COORD])
        A = env[pc : pc + 3]
        es = env[int(bas[ib, PTR_EXP]) : int(bas[ib, PTR_EXP]) + nprim]
        cs = env[int(bas[ib, PTR_COEFF]) : int(bas[ib, PTR_COEFF]) + nprim * nctr].reshape(nctr, nprim)
        x = coords[:, 0] - A[0]
        y = coords[:, 1] - A[1]
        z = coords[:, 2] - A[2]
        rr = x * x + y * y + z * z
        e = np.exp(-np.outer(rr, es))  # [G, nprim]
        for ic in range(nctr):
            R0 = e @ cs[ic]  # sum_p c_p exp(-a r^2)
            R1 = e @ (cs[ic] * (-2.0 * es))  # (1/x) dR/dx
            for m, terms in enumerate(_SOLID[l]):
                val, gx, gy, gz = _poly_and_grad(terms, x, y, z)
                out[0, :, i0 + m] = val * R0
                if deriv == 1:
                    out[1, :, i0 + m] = gx * R0 + val * x * R1
                    out[2, :, i0 + m] = gy * R0 + val * y * R1
                    out[3, :, i0 + m] = gz * R0 + val * z * R1
            i0 += 2 * l + 1
    return out[0] if deriv == 0 else out
