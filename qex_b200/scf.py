"""GPU-resident fixed-shape SCF iteration around the hot path.

Host mirror (torch, float64, everything stays on the device) of the reference's jit-shaped prototype:
    scf_functions_masked.py:143-193   get_veff_jax / energy_tot_jax / get_occ / make_rdm1
    scf_functions_masked.py:856-904   _scf_test_non_padded  (the loop; `scf_loop` here)
    jax_diis.py:17-132                initialize_diis / update_diis_state / extrapolate_fock / apply_diis
    generalized_eigensolver.py:264-330  generalized_eigh      (Cholesky + two triangular solves + eigh)
    generalized_eigensolver.py:143-219  degen_eigh + degen_eigh_bwd (degenerate-safe cotangent)
    hf_legacy.py get_veff / energy_elec  (`rhf_loop`: vj - vk/2)

What runs where: the two heavy pieces of every cycle are this repo's CUDA kernels --
V_xc/E_xc through `autograd.nr_rks` (csrc/contract.cu, xc_*.cu) and J/K through `hf`
(csrc/jk.cu); the N x N linear algebra (Cholesky, triangular solves, eigh, the DIIS solve) is
library code (torch.linalg -> cuSOLVER/cuBLAS), exactly the role jnp.linalg plays in the reference.
The loop is differentiable end to end with torch.autograd (d e_tot / d theta through all cycles),
which is what the reference's trainer does with jax.grad.
"""
from __future__ import annotations

import torch

from . import autograd as _ag
from . import hf


# ---------------------------------------------------------------- DIIS (jax_diis.py)
def initialize_diis(max_vec: int = 6) -> dict:
    return dict(error_vecs=[], fock_vecs=[], B=None, iteration=0, max_vec=max_vec)


def get_diis_error(fock, dm, ovlp):
    return fock @ (dm @ ovlp) - ovlp @ (dm @ fock)


def update_diis_state(state: dict, error_vec, fock, max_vec: int = 6) -> dict:
    ev = list(state["error_vecs"]) + [error_vec.reshape(-1)]
    fv = list(state["fock_vecs"]) + [fock.reshape(-1)]
    if len(ev) > max_vec:
        ev, fv = ev[-max_vec:], fv[-max_vec:]
    n = len(ev)
    E = torch.stack(ev)
    B = torch.zeros(n + 1, n + 1, dtype=fock.dtype, device=fock.device)
    B[0, 1:] = -1.0  # "The -1 is critical here" (jax_diis.py:48)
    B[1:, 0] = -1.0
    B[1:, 1:] = E @ E.T
    return dict(error_vecs=ev, fock_vecs=fv, B=B, iteration=state["iteration"] + 1, max_vec=max_vec)


def extrapolate_fock(state: dict, fock_shape, min_vecs: int = 2, damping: float = 0.0):
    n = len(state["fock_vecs"])
    if n < min_vecs:
        if n > 0:
            return state["fock_vecs"][-1].reshape(fock_shape)
        raise ValueError("no Fock matrix stored yet")
    B = state["B"]
    rhs = torch.zeros(n + 1, dtype=B.dtype, device=B.device)
    rhs[0] = -1.0
    B_reg = B + 1e-14 * torch.eye(n + 1, dtype=B.dtype, device=B.device)
    c = torch.linalg.solve(B_reg, rhs)
    f = (c[1:, None] * torch.stack(state["fock_vecs"])).sum(0)
    if damping > 0.0:
        f = (1.0 - damping) * f + damping * state["fock_vecs"][-1]
    return f.reshape(fock_shape)


def apply_diis(state: dict, fock, dm, ovlp, max_vec: int = 6, min_vecs: int = 2, damping: float = 0.0):
    new = update_diis_state(state, get_diis_error(fock, dm, ovlp), fock, max_vec)
    return extrapolate_fock(new, fock.shape, min_vecs, damping), new


# ---------------------------------------------------------------- eigensolver (row N3)
class _DegenEigh(torch.autograd.Function):
    """`degen_eigh` with the reference's custom cotangent: 1/(l_j - l_i) is dropped (set to 0) for
    |l_j - l_i| < eps**0.6, eigenvalue term V diag(g) V^T, result symmetrised."""

    @staticmethod
    def forward(ctx, A):
        w, V = torch.linalg.eigh(A)
        ctx.save_for_backward(w, V)
        return w, V

    @staticmethod
    def backward(ctx, gw, gV):
        w, V = ctx.saved_tensors
        Vt = V.transpose(-1, -2)
        res = torch.zeros_like(V)
        if gV is not None:
            thr = torch.finfo(w.dtype).eps ** 0.6
            F = w.unsqueeze(-2) - w.unsqueeze(-1)
            near = F.abs() < thr
            Finv = torch.where(near, torch.zeros_like(F), 1.0 / torch.where(near, torch.ones_like(F), F))
            res = V @ ((Finv * (Vt @ gV)) @ Vt)
        if gw is not None:
            res = res + V @ (gw.unsqueeze(-1) * Vt)
        return (res + res.transpose(-1, -2)) * 0.5


def degen_eigh(A):
    return _DegenEigh.apply(A)


def generalized_eigh(A, B, eps: float = 1.0e-12, scale: bool = False, degenerate_safe: bool = True):
    """A v = w B v for symmetric A and SPD B (generalized_eigensolver.py:264-330).  `degenerate_safe`
    routes the inner eigh through `degen_eigh` (the rule the reference's masked solver uses)."""
    A = (A + A.transpose(-1, -2)) * 0.5
    B = (B + B.transpose(-1, -2)) * 0.5
    if scale:
        s_inv = 1.0 / torch.sqrt(torch.diagonal(B, dim1=-2, dim2=-1))
        A = (s_inv.unsqueeze(-1) * A) * s_inv.unsqueeze(-2)
        B = (s_inv.unsqueeze(-1) * B) * s_inv.unsqueeze(-2)
    lam_min = torch.linalg.eigvalsh(B.detach()).min()
    shift = torch.clamp(eps - lam_min, min=0.0)
    B = B + shift * torch.eye(B.shape[-1], dtype=B.dtype, device=B.device)
    L = torch.linalg.cholesky(B)
    Y = torch.linalg.solve_triangular(L, A, upper=False)
    C = torch.linalg.solve_triangular(L, Y.transpose(-1, -2), upper=False).transpose(-1, -2)
    C = (C + C.transpose(-1, -2)) * 0.5
    w, U = degen_eigh(C) if degenerate_safe else torch.linalg.eigh(C)
    V = torch.linalg.solve_triangular(L.transpose(-1, -2), U, upper=True)
    return w, V


# ---------------------------------------------------------------- SCF pieces
def get_occ(nelectron: int, mo_energy):
    e_idx = torch.argsort(mo_energy)
    idx = torch.arange(mo_energy.shape[0], device=mo_energy.device)
    mo_occ = torch.where(idx < nelectron // 2, 2.0, 0.0).to(mo_energy.dtype)
    return mo_occ[torch.argsort(e_idx)]


def make_rdm1(mo_coeff, mo_occ):
    return (mo_coeff * mo_occ.unsqueeze(-2)) @ mo_coeff.transpose(-1, -2)


def get_j(eri, dm):
    """J_ij = sum_kl eri[i,j,k,l] dm[k,l]  (`einsum("ijkl,kl->ij")`, scf_functions_masked.py:152) on the
    streaming kernel: it is the J-bar half of the transposed contraction, read back transposed."""
    return hf.dot_eri_dm_rowdot(eri, dm)


def _nr_rks_lda(xc, dm):
    """`nr_rks(mol, grids, "lda", dm)` on the device without a network: rho (K2) -> Slater exchange (csrc/xc_lda.cu)
    -> E_xc / V_xc assembly (K5).  Not differentiable (there is nothing to train)."""
    from .xc import lda_exchange

    with torch.no_grad():
        B, N = xc.nbatch, xc.nao
        rho = xc.eval_rho(dm.detach().reshape(B, N, N), 1, 1)
        exc, vrho = lda_exchange(rho[:, 0, :])
        out = xc.vxc_assemble(rho, exc, vrho, None, "NN")
    return out[:, N * N + 1], out[:, N * N], out[:, : N * N].reshape(B, N, N)


def get_veff(xc, dm, eri, theta, xctype: str = "NN"):
    """-> (J + V_xc, E_xc, J) of get_veff_jax; the grid, weights and AO values live in the XCContext.
    ``theta=None`` with ``xctype="LDA"`` selects the analytic Slater-exchange functional (pyscf's ``xc = "lda"``)."""
    J = get_j(eri, dm)
    if theta is None and xctype == "LDA":
        _nelec, excsum, vmat = _nr_rks_lda(xc, dm)
    else:
        _nelec, excsum, vmat = _ag.nr_rks(xc, dm, theta, xctype, hermi=1)
    return J + vmat[0], excsum[0], J


def energy_tot(dm, h1e, J, exc_energy, energy_nuc):
    return (dm * h1e.T).sum() + 0.5 * (dm * J).sum() + exc_energy + energy_nuc


def scf_loop(xc, theta, dm, eri, s1e, h1e, energy_nuc, nelectron, xctype="NN", max_cycle=15, diis_max_vec=15,
             diis_min_vec=2, diis_start_cycle=1, diis_damping=0.0):
    """`_scf_test_non_padded`: returns (e_tot, dm, energies[max_cycle]); differentiable w.r.t. theta."""
    vhf, exc_e, J = get_veff(xc, dm, eri, theta, xctype)
    e_tot = energy_tot(dm, h1e, J, exc_e, energy_nuc)
    st = initialize_diis(diis_max_vec)
    energies = []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            fock, st = apply_diis(st, fock, dm, s1e, diis_max_vec, diis_min_vec, diis_damping)
        mo_energy, mo_coeff = generalized_eigh(fock, s1e)
        dm = make_rdm1(mo_coeff, get_occ(nelectron, mo_energy))
        vhf, exc_e, J = get_veff(xc, dm, eri, theta, xctype)
        e_tot = energy_tot(dm, h1e, J, exc_e, energy_nuc)
        energies.append(e_tot)
    return e_tot, dm, torch.stack(energies)


def rhf_loop(dm, eri, s1e, h1e, energy_nuc, nelectron, max_cycle=30, diis_max_vec=15, diis_min_vec=2,
             diis_start_cycle=1):
    """The same loop with the Hartree-Fock potential vj - vk/2 (J and K from one pass over the tensor)."""
    def veff(d):
        vj, vk = hf.dot_eri_dm_autograd(eri, d)
        return vj - 0.5 * vk

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
        energies.append((h1e * dm.T).sum() + 0.5 * (vhf * dm.T).sum() + energy_nuc)
    return energies[-1], dm, torch.stack(energies)


def core_guess(h1e, s1e, nelectron):
    e, c = generalized_eigh(h1e, s1e)
    return make_rdm1(c, get_occ(nelectron, e))


# ---------------------------------------------------------------- implicit differentiation (row N4)
def scf_optimality_cond(xc, theta, dm, eri, s1e, h1e, nelectron, xctype="NN"):
    """One application of the SCF map dm -> dm' without DIIS (`_scf_optimality_cond`, hf_legacy.py:40-48)."""
    vhf, _, _ = get_veff(xc, dm, eri, theta, xctype)
    mo_energy, mo_coeff = generalized_eigh(h1e + vhf, s1e)
    return make_rdm1(mo_coeff, get_occ(nelectron, mo_energy))


def _gmres(matvec, b, tol=1e-12, max_iter=50):
    """Plain (unrestarted) GMRES with modified Gram-Schmidt for the small adjoint system (I - J^T) x = b;
    vectors live on the device, the (max_iter+1) x max_iter Hessenberg least-squares problem on the host."""
    import numpy as np

    bnorm = float(b.norm())
    if bnorm == 0.0:
        return torch.zeros_like(b)
    Q = [b / bnorm]
    Hm = np.zeros((max_iter + 1, max_iter))
    y = None
    for k in range(max_iter):
        w = matvec(Q[k])
        for i in range(k + 1):
            Hm[i, k] = float((Q[i] * w).sum())
            w = w - Hm[i, k] * Q[i]
        Hm[k + 1, k] = float(w.norm())
        e1 = np.zeros(k + 2)
        e1[0] = bnorm
        y, res, *_ = np.linalg.lstsq(Hm[: k + 2, : k + 1], e1, rcond=None)
        r = np.linalg.norm(Hm[: k + 2, : k + 1] @ y - e1)
        if r <= tol * bnorm or Hm[k + 1, k] <= 1e-300:
            break
        Q.append(w / Hm[k + 1, k])
    x = torch.zeros_like(b)
    for i, yi in enumerate(y):
        x = x + float(yi) * Q[i]
    return x


class _ImplicitSCF(torch.autograd.Function):
    """Converged density matrix with the gradient of the FIXED POINT dm* = T(dm*, theta), not of the
    iteration history (`make_implicit_diff(_scf, ..., optimality_cond=_scf_optimality_cond,
    solver=gen_gmres())`, hf_legacy.py:202-205): backward solves (I - dT/ddm)^T lam = dm_bar by GMRES,
    every operator application being one VJP of the XC + J kernels, then theta_bar = (dT/dtheta)^T lam."""

    @staticmethod
    def forward(ctx, xc, theta, dm0, eri, s1e, h1e, nelectron, xctype, max_cycle, conv_tol, diis_max_vec):
        with torch.no_grad():
            dm = dm0
            st = initialize_diis(diis_max_vec)
            for cycle in range(max_cycle):
                vhf, _, _ = get_veff(xc, dm, eri, theta, xctype)
                fock = h1e + vhf
                if cycle >= 1:
                    fock, st = apply_diis(st, fock, dm, s1e, diis_max_vec, 2, 0.0)
                mo_energy, mo_coeff = generalized_eigh(fock, s1e)
                new = make_rdm1(mo_coeff, get_occ(nelectron, mo_energy))
                delta = float((new - dm).norm())
                dm = new
                if delta < conv_tol:
                    break
            # polish with plain applications of T so that dm is a fixed point of T itself
            for _ in range(3):
                dm = scf_optimality_cond(xc, theta, dm, eri, s1e, h1e, nelectron, xctype)
        ctx.args = (xc, eri, s1e, h1e, nelectron, xctype)
        ctx.save_for_backward(theta.detach(), dm)
        return dm.clone()

    @staticmethod
    def backward(ctx, dm_bar):
        xc, eri, s1e, h1e, nelectron, xctype = ctx.args
        theta, dm = ctx.saved_tensors
        with torch.enable_grad():
            th = theta.detach().requires_grad_(True)
            d = dm.detach().requires_grad_(True)
            out = scf_optimality_cond(xc, th, d, eri, s1e, h1e, nelectron, xctype)

            def jt(v):  # (dT/ddm)^T v
                (g,) = torch.autograd.grad(out, d, v, retain_graph=True)
                return g

            lam = _gmres(lambda v: v - jt(v), dm_bar.contiguous())
            (theta_bar,) = torch.autograd.grad(out, th, lam, retain_graph=False)
        return None, theta_bar, None, None, None, None, None, None, None, None, None


def scf_fixed_point(xc, theta, dm0, eri, s1e, h1e, nelectron, xctype="NN", max_cycle=50, conv_tol=1e-11,
                    diis_max_vec=8):
    """Self-consistent density matrix, differentiable w.r.t. theta by implicit differentiation."""
    return _ImplicitSCF.apply(xc, theta, dm0, eri, s1e, h1e, nelectron, xctype, max_cycle, conv_tol, diis_max_vec)


# ---------------------------------------------------------------- batched over molecules (row N1, config c4)
def _diis_batched(ev, fv, fock_shape, min_vecs):
    n = len(fv)
    if n < min_vecs:
        return fv[-1].reshape(fock_shape)
    E = torch.stack(ev, dim=1)  # [B, n, N*N]
    nb = E.shape[0]
    Bm = torch.zeros(nb, n + 1, n + 1, dtype=E.dtype, device=E.device)
    Bm[:, 0, 1:] = -1.0
    Bm[:, 1:, 0] = -1.0
    Bm[:, 1:, 1:] = E @ E.transpose(-1, -2)
    rhs = torch.zeros(nb, n + 1, dtype=E.dtype, device=E.device)
    rhs[:, 0] = -1.0
    # solve_ex without the host-side `info` check: the loop stays asynchronous (and CUDA-graph capturable)
    c, _ = torch.linalg.solve_ex(Bm + 1e-14 * torch.eye(n + 1, dtype=E.dtype, device=E.device), rhs, check_errors=False)
    return (c[:, 1:, None] * torch.stack(fv, dim=1)).sum(1).reshape(fock_shape)


def get_occ_batched(nelectron: int, mo_energy):
    ranks = torch.argsort(torch.argsort(mo_energy, dim=-1), dim=-1)
    return torch.where(ranks < nelectron // 2, 2.0, 0.0).to(mo_energy.dtype)


class _GEighSmall(torch.autograd.Function):
    """`generalized_eigh` for stacks of small matrices (n <= 16) on the one-thread-per-matrix Jacobi kernel
    (csrc/eigh.cu): no cuSOLVER launch chain and no host synchronisation per SCF cycle.  Backward is the
    reference's degenerate-safe rule in the B-orthonormal basis, A_bar = sym(V (diag(w_bar) + F o (V^T V_bar)) V^T);
    B (the overlap matrix) is treated as a constant, as it is in the SCF loop."""

    @staticmethod
    def forward(ctx, A, Bm, eps):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        nb, n = int(A.shape[0]), int(A.shape[-1])
        A = A.contiguous()
        Bm = Bm.contiguous()
        w = torch.empty(nb, n, dtype=torch.float64, device=A.device)
        V = torch.empty(nb, n, n, dtype=torch.float64, device=A.device)
        with torch.cuda.device(A.device):
            _lib.check(lib.qexxc_generalized_eigh_batched(A.device.index, A.data_ptr(), Bm.data_ptr(), nb, n, float(eps),
                                                          w.data_ptr(), V.data_ptr(),
                                                          C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)))
        ctx.save_for_backward(w, V)
        return w, V

    @staticmethod
    def backward(ctx, gw, gV):
        w, V = ctx.saved_tensors
        Vt = V.transpose(-1, -2)
        inner = torch.zeros_like(V)
        if gV is not None:
            thr = torch.finfo(w.dtype).eps ** 0.6
            F = w.unsqueeze(-2) - w.unsqueeze(-1)
            near = F.abs() < thr
            Finv = torch.where(near, torch.zeros_like(F), 1.0 / torch.where(near, torch.ones_like(F), F))
            inner = Finv * (Vt @ gV)
        if gw is not None:
            inner = inner + torch.diag_embed(gw)
        res = V @ inner @ Vt
        return (res + res.transpose(-1, -2)) * 0.5, None, None


def generalized_eigh_batched(A, Bm, eps: float = 1.0e-12, small_kernel: bool | None = None):
    """`generalized_eigh` for stacks [B, N, N] (per-matrix SPD shift).  N <= 16 on a CUDA device goes to the
    batched Jacobi kernel unless `small_kernel=False`; otherwise torch.linalg (cuSOLVER)."""
    if small_kernel is None:
        small_kernel = A.is_cuda and A.shape[-1] <= 16 and not Bm.requires_grad
    if small_kernel:
        return _GEighSmall.apply(A, Bm, eps)
    A = (A + A.transpose(-1, -2)) * 0.5
    Bm = (Bm + Bm.transpose(-1, -2)) * 0.5
    lam_min = torch.linalg.eigvalsh(Bm.detach()).amin(-1)
    shift = torch.clamp(eps - lam_min, min=0.0)
    Bm = Bm + shift[:, None, None] * torch.eye(Bm.shape[-1], dtype=Bm.dtype, device=Bm.device)
    L = torch.linalg.cholesky(Bm)
    Y = torch.linalg.solve_triangular(L, A, upper=False)
    Cm = torch.linalg.solve_triangular(L, Y.transpose(-1, -2), upper=False).transpose(-1, -2)
    Cm = (Cm + Cm.transpose(-1, -2)) * 0.5
    w, U = degen_eigh(Cm)
    return w, torch.linalg.solve_triangular(L.transpose(-1, -2), U, upper=True)


def get_veff_batched(xc, dm, eri, theta, xctype: str = "NN"):
    J = hf.dot_eri_dm_rowdot_batched(eri, dm)
    if theta is None and xctype == "LDA":
        _nelec, excsum, vmat = _nr_rks_lda(xc, dm)
    else:
        _nelec, excsum, vmat = _ag.nr_rks(xc, dm, theta, xctype, hermi=1)
    return J + vmat, excsum, J


def scf_loop_batched(xc, theta, dm, eri, s1e, h1e, energy_nuc, nelectron, xctype="NN", max_cycle=15, diis_max_vec=15,
                     diis_min_vec=2, diis_start_cycle=1):
    """`scf_loop` for B molecules of equal nao at once (XCContext with nbatch = B): every cycle is one batched
    XC launch sequence, one batched J launch and batched N x N linear algebra.  dm, s1e, h1e: [B, N, N];
    eri: [B, N, N, N, N]; energy_nuc: [B].  -> (e_tot [B], dm [B, N, N], energies [max_cycle, B])."""
    def energy(d, J, exc_e):
        return (d * h1e.transpose(-1, -2)).sum((-1, -2)) + 0.5 * (d * J).sum((-1, -2)) + exc_e + energy_nuc

    vhf, exc_e, J = get_veff_batched(xc, dm, eri, theta, xctype)
    e_tot = energy(dm, J, exc_e)
    ev, fv, energies = [], [], []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            ev = (ev + [get_diis_error(fock, dm, s1e).reshape(fock.shape[0], -1)])[-diis_max_vec:]
            fv = (fv + [fock.reshape(fock.shape[0], -1)])[-diis_max_vec:]
            fock = _diis_batched(ev, fv, fock.shape, diis_min_vec)
        mo_energy, mo_coeff = generalized_eigh_batched(fock, s1e)
        dm = make_rdm1(mo_coeff, get_occ_batched(nelectron, mo_energy))
        vhf, exc_e, J = get_veff_batched(xc, dm, eri, theta, xctype)
        e_tot = energy(dm, J, exc_e)
        energies.append(e_tot)
    return e_tot, dm, torch.stack(energies)


# ---------------------------------------------------------------- padded / masked batches (row N1, second half)
# Mixed-size batches (different nao and grid sizes per molecule) run through the same fixed-shape kernels once every
# molecule is padded to the batch maximum: zero rows/columns in dm, h1e, s1e and eri, zero AO columns for padded
# orbitals, zero weights (and zero AO rows) for padded grid points.  With that padding stages 2-4 and the J kernel need
# no mask at all (padded entries contribute exact zeros); the mask matters in the eigensolver, the occupations and the
# density-matrix build.  Mirrors scf_functions_masked.py:244-309,546-588,917-967 and
# generalized_eigensolver_masked.py:19-89.  The reference's "stable" get_veff pads with eps = 1e-12 instead of 0 and
# adds eps to rho at every grid point (scf_functions_masked.py:567-579); over a grid that reaches ~100 Bohr that moves
# E_xc by ~1e-6 Ha (measured with the oracle, tests/test_scf_masked.py) -- a prototype artefact, not reproduced: this
# loop equals the unpadded one to 1e-10 Ha.
def pad_stack(mats, n: int | None = None):
    """List of square [N_b, N_b] (or [N_b]^4) tensors -> zero-padded stack [B, n, n(, n, n)] and mask [B, n]."""
    n = n or max(int(m.shape[0]) for m in mats)
    out = torch.zeros((len(mats),) + (n,) * mats[0].dim(), dtype=mats[0].dtype, device=mats[0].device)
    mask = torch.zeros(len(mats), n, dtype=torch.bool, device=mats[0].device)
    for b, m in enumerate(mats):
        k = int(m.shape[0])
        out[(b,) + (slice(0, k),) * m.dim()] = m
        mask[b, :k] = True
    return out, mask


def masked_generalized_eigh(fock, s1e, mask, eps: float = 1.0e-12):
    """`masked_generalized_eigh` (generalized_eigensolver_masked.py:19-89) for [..., N, N] stacks: padded block of the
    Fock matrix zeroed with 1e-12 on its diagonal, padded diagonal of the overlap set to 1e-12, degenerate-safe
    generalised eigensolver, real eigenpairs first in ascending order (the reference keys the sort on the mask by
    POSITION, reproduced here), padded eigenvalues and coefficient rows/columns zeroed."""
    single = fock.dim() == 2
    if single:
        fock, s1e, mask = fock[None], s1e[None], mask[None]
    n = fock.shape[-1]
    m2 = mask[:, :, None] & mask[:, None, :]
    eye = torch.eye(n, dtype=torch.bool, device=fock.device)
    pad_diag = torch.where((~mask)[:, :, None] & eye, torch.full_like(fock, 1e-12), torch.zeros_like(fock))
    f = torch.where(m2, fock, torch.zeros_like(fock)) + pad_diag
    s = torch.where(m2, s1e, pad_diag)
    w, v = generalized_eigh_batched(f, s, eps)
    key = torch.where(mask, w, torch.full_like(w, 1e12))
    idx = torch.argsort(key, dim=-1, stable=True)
    w = torch.gather(w, -1, idx) * torch.gather(mask, -1, idx).to(w.dtype)
    v = torch.gather(v, -1, idx[:, None, :].expand_as(v))
    keep = mask[:, :, None] & torch.gather(mask, -1, idx)[:, None, :]
    v = torch.where(keep, v, torch.zeros_like(v))
    return (w[0], v[0]) if single else (w, v)


def get_occ_masked(nelectron, mo_energy, mask):
    """`get_occ_masked` (scf_functions_masked.py:269-288); `nelectron` may be an int or a per-molecule tensor [B]."""
    e = torch.where(mask, mo_energy, torch.full_like(mo_energy, 1e10))
    e_idx = torch.argsort(e, dim=-1, stable=True)
    nocc = torch.as_tensor(nelectron, device=mo_energy.device) // 2
    idx = torch.arange(mo_energy.shape[-1], device=mo_energy.device)
    occ_sorted = torch.where((idx < nocc[..., None]) & torch.gather(mask, -1, e_idx), 2.0, 0.0).to(mo_energy.dtype)
    return torch.gather(occ_sorted, -1, torch.argsort(e_idx, dim=-1))


def make_rdm1_masked(mo_coeff, mo_occ, mask):
    """`make_rdm1_masked` (scf_functions_masked.py:524-541)."""
    c = torch.where(mask[..., :, None], mo_coeff, torch.zeros_like(mo_coeff))
    occ = torch.where(mask, mo_occ, torch.zeros_like(mo_occ))
    dm = (c * occ.unsqueeze(-2)) @ c.transpose(-1, -2)
    return torch.where(mask[..., :, None] & mask[..., None, :], dm, torch.zeros_like(dm))


def get_veff_masked(xc, dm, eri, theta, mask, xctype: str = "NN"):
    """`get_veff_jax_masked`: J + V_xc, E_xc, J on zero-padded inputs (the XCContext holds the padded AO values and
    weights); outputs masked again as the reference does."""
    m2 = (mask[..., :, None] & mask[..., None, :]).to(dm.dtype)
    vhf, excsum, J = get_veff_batched(xc, dm * m2, eri, theta, xctype)
    return vhf * m2, excsum, J * m2


def energy_tot_masked(dm, h1e, J, exc_energy, energy_nuc, mask):
    """`energy_tot_jax_masked` (scf_functions_masked.py:244-265), batched."""
    m2 = (mask[..., :, None] & mask[..., None, :]).to(dm.dtype)
    dm, h1e, J = dm * m2, h1e * m2, J * m2
    return (dm * h1e.transpose(-1, -2)).sum((-1, -2)) + 0.5 * (dm * J).sum((-1, -2)) + exc_energy + energy_nuc


def scf_loop_padded(xc, theta, dm, eri, s1e, h1e, energy_nuc, nelectron, mask, xctype="NN", max_cycle=15,
                    diis_max_vec=15, diis_min_vec=2, diis_start_cycle=1):
    """`_scf_test_padded` (scf_functions_masked.py:917-967) for a batch of molecules of DIFFERENT sizes padded to a
    common [B, N, N] (see `pad_stack`); nelectron: int or [B].  -> (e_tot [B], dm [B, N, N], energies [cycles, B])."""
    vhf, exc_e, J = get_veff_masked(xc, dm, eri, theta, mask, xctype)
    e_tot = energy_tot_masked(dm, h1e, J, exc_e, energy_nuc, mask)
    ev, fv, energies = [], [], []
    for cycle in range(max_cycle):
        fock = h1e + vhf
        if cycle >= diis_start_cycle:
            ev = (ev + [get_diis_error(fock, dm, s1e).reshape(fock.shape[0], -1)])[-diis_max_vec:]
            fv = (fv + [fock.reshape(fock.shape[0], -1)])[-diis_max_vec:]
            fock = _diis_batched(ev, fv, fock.shape, diis_min_vec)
        mo_energy, mo_coeff = masked_generalized_eigh(fock, s1e, mask)
        dm = make_rdm1_masked(mo_coeff, get_occ_masked(nelectron, mo_energy, mask), mask)
        vhf, exc_e, J = get_veff_masked(xc, dm, eri, theta, mask, xctype)
        e_tot = energy_tot_masked(dm, h1e, J, exc_e, energy_nuc, mask)
        energies.append(e_tot)
    return e_tot, dm, torch.stack(energies)
