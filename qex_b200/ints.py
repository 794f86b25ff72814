"""Host-side one- and two-electron integrals over contracted s-type Gaussians (closed form, Boys F0 via
erf), from libcint tables -- an INPUT GENERATOR for workloads and examples (H2 / 6-31G dissociation curves,
config c4), like gen_grid.py and gto.py.  The reference takes these arrays from pyscf
(`mol.intor("int1e_ovlp" / "int1e_kin" / "int1e_nuc" / "int2e")`, hf_legacy.py:145-154); shells with l > 0
raise NotImplementedError.  tests/test_scf.py checks this module against oracle/ints_ref.py and, through the
RHF loop, against the six energies printed by the reference's notebook."""
import math

import numpy as np

try:
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, PTR_EXP, PTR_COEFF = 0, 1, 2, 3, 5, 6
CHARGE_OF, PTR_COORD = 0, 1


def _boys0(t):
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-12
    ts = np.where(small, 1.0, t)
    val = 0.5 * np.sqrt(np.pi / ts) * _erf(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0, val)


def _primitives(atm, bas, env):
    """Flatten to primitive s Gaussians c * exp(-a |r-A|^2): arrays (ao index, a, c, centre).
    libcint stores radial-normalised coefficients; the real solid harmonic Y00 = 1/sqrt(4 pi)
    completes the AO (same convention as oracle/gto_ref.py)."""
    idx, a, c, A = [], [], [], []
    ao = 0
    for b in bas:
        if b[ANG_OF] != 0:
            raise NotImplementedError("closed-form integrals are restated for s shells only")
        nprim, nctr = int(b[NPRIM_OF]), int(b[NCTR_OF])
        ex = env[b[PTR_EXP]: b[PTR_EXP] + nprim]
        cf = env[b[PTR_COEFF]: b[PTR_COEFF] + nprim * nctr].reshape(nctr, nprim)
        xyz = env[atm[b[ATOM_OF], PTR_COORD]: atm[b[ATOM_OF], PTR_COORD] + 3]
        for k in range(nctr):
            for p in range(nprim):
                idx.append(ao)
                a.append(ex[p])
                c.append(cf[k, p] / math.sqrt(4.0 * math.pi))
                A.append(xyz)
            ao += 1
    return np.array(idx), np.array(a), np.array(c), np.array(A), ao


def integrals(atm, bas, env):
    """-> dict(s1e, t1e, v1e, h1e [N,N], eri [N,N,N,N] (chemists' (ij|kl), s1), enuc)."""
    atm, bas, env = np.asarray(atm), np.asarray(bas), np.asarray(env, dtype=np.float64)
    idx, a, c, A, nao = _primitives(atm, bas, env)
    npr = len(a)
    p = a[:, None] + a[None, :]
    mu = a[:, None] * a[None, :] / p
    R2 = ((A[:, None, :] - A[None, :, :]) ** 2).sum(-1)
    K = np.exp(-mu * R2)
    cc = c[:, None] * c[None, :]
    S = cc * (np.pi / p) ** 1.5 * K
    T = mu * (3.0 - 2.0 * mu * R2) * S
    P = (a[:, None, None] * A[:, None, :] + a[None, :, None] * A[None, :, :]) / p[:, :, None]
    V = np.zeros_like(S)
    for at in atm:
        Z = float(at[CHARGE_OF])
        C = env[at[PTR_COORD]: at[PTR_COORD] + 3]
        V += -Z * cc * (2.0 * np.pi / p) * K * _boys0(p * ((P - C) ** 2).sum(-1))
    # (ab|cd) over primitives
    pq = p[:, :, None, None] * p[None, None, :, :]
    ps = p[:, :, None, None] + p[None, None, :, :]
    PQ2 = ((P[:, :, None, None, :] - P[None, None, :, :, :]) ** 2).sum(-1)
    E = (2.0 * np.pi ** 2.5 / (pq * np.sqrt(ps)) * (cc * K)[:, :, None, None] * (cc * K)[None, None, :, :]
         * _boys0(pq / ps * PQ2))
    # contract primitives -> AOs
    M = np.zeros((nao, npr))
    M[idx, np.arange(npr)] = 1.0
    s1e, t1e, v1e = (M @ X @ M.T for X in (S, T, V))
    eri = np.einsum("ia,jb,kc,ld,abcd->ijkl", M, M, M, M, E, optimize=True)
    enuc = 0.0
    for i in range(len(atm)):
        for j in range(i):
            ri = env[atm[i, PTR_COORD]: atm[i, PTR_COORD] + 3]
            rj = env[atm[j, PTR_COORD]: atm[j, PTR_COORD] + 3]
            enuc += float(atm[i, CHARGE_OF]) * float(atm[j, CHARGE_OF]) / float(np.linalg.norm(ri - rj))
    return dict(s1e=s1e, t1e=t1e, v1e=v1e, h1e=t1e + v1e, eri=eri, enuc=enuc)
