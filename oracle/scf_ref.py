"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's
fixed-shape SCF iteration in numpy/scipy.

Follows, function by function:
  jax_diis.py:17-132              initialize_diis / update_diis_state / extrapolate_fock /
                                  get_diis_error / apply_diis  (B[0,1:] = B[1:,0] = -1, rhs[0] = -1,
                                  1e-14 on the diagonal, coefficients c[1:])
  generalized_eigensolver.py:264-330  generalized_eigh (symmetrise, optional Jacobi scaling, SPD shift,
                                  Cholesky, two triangular solves, eigh, back-transform)
  generalized_eigensolver.py:143-219  degen_eigh_bwd (degenerate-safe eigh cotangent)
  scf_functions_masked.py:143-193 get_veff_jax / energy_tot_jax / get_occ / make_rdm1
  scf_functions_masked.py:856-904 _scf_test_non_padded (the loop)
  hf_legacy.py / rks_legacy.py    RHF effective potential vj - vk/2 and E = sum(h dm) + sum(vhf dm)/2 + E_nuc
Pinned: the RHF loop reproduces the six H2/6-31G energies frozen in the reference's notebook
(tests/test_scf.py), which pins the J/K contraction, the eigensolver and the loop together.
"""
import numpy as np
import scipy.linalg as sla

from . import jk_ref


# ---------------------------------------------------------------- DIIS (jax_diis.py)
def initialize_diis(max_vec=6):
    return dict(error_vecs=[], fock_vecs=[], B=np.zeros((max_vec + 1, max_vec + 1)), iteration=0)


def get_diis_error(fock, dm, ovlp):
    return fock @ (dm @ ovlp) - ovlp @ (dm @ fock)


def update_diis_state(state, error_vec, fock, max_vec=6):
    ev = list(state["error_vecs"]) + [error_vec.ravel()]
    fv = list(state["fock_vecs"]) + [fock.ravel()]
    if len(ev) > max_vec:
        ev, fv = ev[-max_vec:], fv[-max_vec:]
    n = len(ev)
    B = np.zeros((n + 1, n + 1))
    B[0, 1:] = -1.0
    B[1:, 0] = -1.0
    for i in range(n):
        for j in range(n):
            B[i + 1, j + 1] = np.dot(ev[i], ev[j])
    return dict(error_vecs=ev, fock_vecs=fv, B=B, iteration=state["iteration"] + 1)


def extrapolate_fock(state, fock_shape, min_vecs=2, damping=0.0):
    n = len(state["fock_vecs"])
    if n < min_vecs:
        return state["fock_vecs"][-1].reshape(fock_shape) if n > 0 else np.zeros(fock_shape)
    rhs = np.zeros(n + 1)
    rhs[0] = -1.0
    B = state["B"][: n + 1, : n + 1].copy()
    B[np.diag_indices(n + 1)] += 1e-14
    c = np.linalg.solve(B, rhs)
    f = np.zeros_like(state["fock_vecs"][0])
    for i, ci in enumerate(c[1:]):
        f = f + ci * state["fock_vecs"][i]
    if damping > 0.0:
        f = (1.0 - damping) * f + damping * state["fock_vecs"][-1]
    return f.reshape(fock_shape)


def apply_diis(state, fock, dm, ovlp, max_vec=6, min_vecs=2, damping=0.0):
    new = update_diis_state(state, get_diis_error(fock, dm, ovlp), fock, max_vec)
    return extrapolate_fock(new, fock.shape, min_vecs, damping), new


# ---------------------------------------------------------------- eigensolver
def generalized_eigh(A, B, eps=1.0e-12, scale=False):
    A = (A + A.T) * 0.5
    B = (B + B.T) * 0.5
    if scale:
        s_inv = 1.0 / np.sqrt(np.diag(B))
        A = (s_inv[:, None] * A) * s_inv[None, :]
        B = (s_inv[:, None] * B) * s_inv[None, :]
    lam_min = np.min(np.linalg.eigvalsh(B))
    shift = eps - lam_min if lam_min < eps else 0.0
    B = B + shift * np.eye(B.shape[-1])
    L = np.linalg.cholesky(B)
    Y = sla.solve_triangular(L, A, lower=True)
    C = sla.solve_triangular(L, Y.T, lower=True).T
    C = (C + C.T) * 0.5
    w, U = np.linalg.eigh(C)
    V = sla.solve_triangular(L.T, U, lower=False)
    return w, V


def degen_eigh_bwd(eival, eivec, grad_eival, grad_eivec):
    """generalized_eigensolver.py:161-219."""
    thr = np.finfo(eival.dtype).eps ** 0.6
    vt = eivec.T
    result = np.zeros_like(eivec)
    if grad_eivec is not None:
        F = eival[None, :] - eival[:, None]
        with np.errstate(divide="ignore"):
            Finv = np.where(np.abs(F) < thr, 0.0, 1.0 / np.where(np.abs(F) < thr, 1.0, F))
        result = eivec @ ((Finv * (vt @ grad_eivec)) @ vt)
    if grad_eival is not None:
        result = result + eivec @ (grad_eival[:, None] * vt)
    return (result + result.T) * 0.5


# ---------------------------------------------------------------- SCF pieces (scf_functions_masked.py)
def get_occ(nelectron, mo_energy):
    e_idx = np.argsort(mo_energy)
    mo_occ = np.where(np.arange(mo_energy.shape[0]) < nelectron // 2, 2.0, 0.0)
    return mo_occ[np.argsort(e_idx)]


def make_rdm1(mo_coeff, mo_occ):
    return np.einsum("ij,j,kj->ik", mo_coeff, mo_occ, mo_coeff)


def get_veff(dm, eri, ao_grid, grid_weights, exc_vrho):
    """:143-159.  exc_vrho(rho) -> (exc [G], vrho [G])."""
    J = np.einsum("ijkl,kl->ij", eri, dm)
    rho = np.einsum("gi,ij,gj->g", ao_grid, dm, ao_grid)
    exc, vrho = exc_vrho(rho)
    Vxc = np.einsum("gi,g,gj->ij", ao_grid, grid_weights * vrho, ao_grid)
    return J + Vxc, float(np.sum(exc * rho * grid_weights)), J


def energy_tot(dm, h1e, J, exc_energy, energy_nuc):
    return float(np.einsum("ij,ji->", dm, h1e) + 0.5 * np.einsum("ij,ij->", dm, J) + exc_energy + energy_nuc)


def scf_loop(dm, eri, ao_grid, grid_weights, s1e, h1e, energy_nuc, nelectron, exc_vrho, max_cycle=15,
             diis_max_vec=15, diis_min_vec=2, diis_start_cycle=1, diis_damping=0.0):
    """_scf_test_non_padded :856-904 -> (e_tot, dm, energies[max_cycle])."""
    vhf, exc_e, J = get_veff(dm, eri, ao_grid, grid_weights, exc_vrho)
    e_tot = energy_tot(dm, h1e, J, exc_e, energy_nuc)
    st = initialize_diis(diis_max_vec)
    energies = []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            fock, st = apply_diis(st, fock, dm, s1e, diis_max_vec, diis_min_vec, diis_damping)
        mo_energy, mo_coeff = generalized_eigh(fock, s1e)
        dm = make_rdm1(mo_coeff, get_occ(nelectron, mo_energy))
        vhf, exc_e, J = get_veff(dm, eri, ao_grid, grid_weights, exc_vrho)
        e_tot = energy_tot(dm, h1e, J, exc_e, energy_nuc)
        energies.append(e_tot)
    return e_tot, dm, np.array(energies)


def rhf_loop(dm, eri, s1e, h1e, energy_nuc, nelectron, max_cycle=30, diis_max_vec=15, diis_min_vec=2,
             diis_start_cycle=1):
    """The same loop with the Hartree-Fock potential vj - vk/2 (hf_legacy.py get_veff / energy_elec)."""
    def veff(d):
        vj, vk = jk_ref.dot_eri_dm(eri, d)
        return vj - 0.5 * vk

    def etot(d, v):
        return float(np.einsum("ij,ji->", h1e, d) + 0.5 * np.einsum("ij,ji->", v, d) + energy_nuc)

    vhf = veff(dm)
    st = initialize_diis(diis_max_vec)
    energies = []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            fock, st = apply_diis(st, fock, dm, s1e, diis_max_vec, diis_min_vec)
        mo_energy, mo_coeff = generalized_eigh(fock, s1e)
        dm = make_rdm1(mo_coeff, get_occ(nelectron, mo_energy))
        vhf = veff(dm)
        energies.append(etot(dm, vhf))
    return energies[-1], dm, np.array(energies)


def core_guess(h1e, s1e, nelectron):
    """dm from the core Hamiltonian (pyscf init_guess='1e'); the converged energy does not depend on it."""
    e, c = generalized_eigh(h1e, s1e)
    return make_rdm1(c, get_occ(nelectron, e))


def scf_fixed_point(dm, eri, ao_grid, grid_weights, s1e, h1e, nelectron, exc_vrho, max_cycle=60, conv_tol=1e-11,
                    diis_max_vec=8):
    """Self-consistent dm = T(dm) of `_scf_optimality_cond` (hf_legacy.py:40-48), reached with DIIS and
    polished with three plain applications of T.  -> (dm, last |T(dm) - dm|_max)."""
    def T(d, fock_hook=None):
        vhf, _, _ = get_veff(d, eri, ao_grid, grid_weights, exc_vrho)
        fock = h1e + vhf
        if fock_hook is not None:
            fock = fock_hook(fock, d)
        e, c = generalized_eigh(fock, s1e)
        return make_rdm1(c, get_occ(nelectron, e))

    st = [initialize_diis(diis_max_vec)]

    def hook(fock, d):
        f, st[0] = apply_diis(st[0], fock, d, s1e, diis_max_vec, 2)
        return f

    for it in range(max_cycle):
        new = T(dm, hook if it >= 1 else None)
        delta = np.abs(new - dm).max()
        dm = new
        if delta < conv_tol:
            break
    for _ in range(3):
        new = T(dm)
        delta = np.abs(new - dm).max()
        dm = new
    return dm, delta


# ---- padded / masked variants (scf_functions_masked.py:244-309,546-588,917-967; generalized_eigensolver_masked.py:19-89)
def masked_generalized_eigh(fock, s1e, mask):
    n = fock.shape[0]
    m2 = mask[:, None] & mask[None, :]
    pad = np.where((~mask)[:, None] & np.eye(n, dtype=bool), 1e-12, 0.0)
    f = np.where(m2, fock, 0.0) + pad
    s = np.where(m2, s1e, pad)
    w, v = generalized_eigh(f, s)
    key = np.where(mask, w, 1e12)  # the reference keys on the mask by position
    idx = np.argsort(key, kind="stable")
    w = w[idx] * mask[idx]
    v = np.where(m2[:, idx], v[:, idx], 0.0)
    return w, v


def get_occ_masked(nelectron, mo_energy, mask):
    e = np.where(mask, mo_energy, 1e10)
    e_idx = np.argsort(e, kind="stable")
    idx = np.arange(mo_energy.shape[0])
    occ = np.where((idx < nelectron // 2) & mask[e_idx], 2.0, 0.0)
    return occ[np.argsort(e_idx)]


def make_rdm1_masked(mo_coeff, mo_occ, mask):
    c = np.where(mask[:, None], mo_coeff, 0.0)
    occ = np.where(mask, mo_occ, 0.0)
    dm = np.einsum("ij,j,kj->ik", c, occ, c)
    return np.where(mask[:, None] & mask[None, :], dm, 0.0)


def get_veff_masked(dm, eri, ao_grid, grid_weights, mask, exc_vrho, eps=1e-12):
    """The definition in force in the reference module (the later, "stable" one): padded entries are eps, rho += eps."""
    m2 = mask[:, None] & mask[None, :]
    m4 = mask[:, None, None, None] & mask[None, :, None, None] & mask[None, None, :, None] & mask[None, None, None, :]
    dm_m = np.where(m2, dm, eps)
    eri_m = np.where(m4, eri, eps)
    ao_m = np.where(mask[None, :], ao_grid, eps)
    J = np.einsum("ijkl,kl->ij", eri_m, dm_m)
    rho = np.einsum("gi,ij,gj->g", ao_m, dm_m, ao_m) + eps
    exc, vrho = exc_vrho(rho)
    vxc = np.einsum("gi,g,gj->ij", ao_m, grid_weights * vrho, ao_m)
    J = np.where(m2, J, 0.0)
    vxc = np.where(m2, vxc, 0.0)
    return J + vxc, float(np.sum(exc * rho * grid_weights)), J


def energy_tot_masked(dm, h1e, J, exc_energy, energy_nuc, mask):
    m2 = mask[:, None] & mask[None, :]
    dm, h1e, J = np.where(m2, dm, 0.0), np.where(m2, h1e, 0.0), np.where(m2, J, 0.0)
    return float(np.einsum("ij,ji->", dm, h1e) + 0.5 * np.einsum("ij,ij->", dm, J) + exc_energy + energy_nuc)


def scf_loop_padded(dm, eri, ao_grid, grid_weights, s1e, h1e, energy_nuc, nelectron, mask, exc_vrho, max_cycle=15,
                    diis_max_vec=15, diis_min_vec=2, diis_start_cycle=1, diis_damping=0.0, eps=1e-12):
    vhf, exc_e, J = get_veff_masked(dm, eri, ao_grid, grid_weights, mask, exc_vrho, eps)
    e_tot = energy_tot_masked(dm, h1e, J, exc_e, energy_nuc, mask)
    st = initialize_diis(diis_max_vec)
    energies = []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            fock, st = apply_diis(st, fock, dm, s1e, diis_max_vec, diis_min_vec, diis_damping)
        mo_energy, mo_coeff = masked_generalized_eigh(fock, s1e, mask)
        dm = make_rdm1_masked(mo_coeff, get_occ_masked(nelectron, mo_energy, mask), mask)
        vhf, exc_e, J = get_veff_masked(dm, eri, ao_grid, grid_weights, mask, exc_vrho, eps)
        e_tot = energy_tot_masked(dm, h1e, J, exc_e, energy_nuc, mask)
        energies.append(e_tot)
    return e_tot, dm, np.array(energies)
