"""SCF iteration around the hot path and the golden values that pin it.

The reference's notebooks freeze six converged RHF energies of H2 / 6-31G (pyscf output in
notebooks/04_notebook_td_trainer.ipynb, cells 1 and 5: train bond lengths 0.74/0.5/1.5 A, validation
0.6/0.9/1.2 A).  They pin, together: the basis tables (qex_b200/gto.py), the closed-form integrals
(oracle/ints_ref.py), the J/K contraction (oracle/jk_ref.py <-> csrc/jk.cu), the generalised
eigensolver, DIIS and the loop (oracle/scf_ref.py <-> qex_b200/scf.py)."""
import numpy as np
import pytest

from oracle import gto_ref, ints_ref, mlp_ref, scf_ref
from qex_b200 import gen_grid, gto

# bond length (Angstrom) -> "converged SCF energy" printed by the reference's notebook
GOLDEN_RHF = {0.74: -1.12675531719693, 0.5: -1.05802481296927, 1.5: -0.997497294328357,
              0.6: -1.11003089523311, 0.9: -1.11168637398406, 1.2: -1.05575928255497}


def _h2(R):
    m = gto.h2(R, "6-31g")
    return m, ints_ref.integrals(m._atm, m._bas, m._env)


def test_integrals_are_sane():
    m, I = _h2(0.74)
    assert np.allclose(np.diag(I["s1e"]), 1.0, atol=1e-12)
    assert np.allclose(I["eri"], I["eri"].transpose(1, 0, 2, 3)) and np.allclose(I["eri"], I["eri"].transpose(2, 3, 0, 1))
    assert abs(I["enuc"] - 1.0 / (0.74 / gto.BOHR)) < 1e-14
    # overlap from the closed form == overlap integrated on a grid with the AO oracle (ties ints_ref to gto_ref)
    g = gen_grid.Grids(m, n_rad=75, n_theta=20, n_phi=20).build()
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    assert np.abs(np.einsum("gi,g,gj->ij", ao, g.weights, ao) - I["s1e"]).max() < 2e-6


@pytest.mark.parametrize("R", list(GOLDEN_RHF))
def test_oracle_rhf_reproduces_the_reference_notebook_energies(R):
    _, I = _h2(R)
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    e, dm, hist = scf_ref.rhf_loop(dm0, I["eri"], I["s1e"], I["h1e"], I["enuc"], 2, max_cycle=30)
    assert abs(e - GOLDEN_RHF[R]) < 5e-12
    assert abs(np.einsum("ij,ji", dm, I["s1e"]) - 2.0) < 1e-12


# bond length -> "E(CCSD)" printed by the same notebook (cell 5); exact for two electrons up to pyscf's CCSD tolerance
GOLDEN_CCSD = {0.74: -1.151672678339737, 0.5: -1.077863888625149, 1.5: -1.054347450987067,
               0.6: -1.131953433438712, 0.9: -1.140602464558199, 1.2: -1.095595490661586}


def _full_ci(I):
    _, C = scf_ref.generalized_eigh(I["h1e"], I["s1e"])
    h = C.T @ I["h1e"] @ C
    e = np.einsum("pi,qj,rk,sl,pqrs->ijkl", C, C, C, C, I["eri"], optimize=True)
    n = h.shape[0]
    eye = np.eye(n)
    H = (np.einsum("ik,jl->ijkl", h, eye) + np.einsum("ik,jl->ijkl", eye, h) + e.transpose(0, 2, 1, 3)).reshape(n * n, n * n)
    return float(np.linalg.eigvalsh(H)[0] + I["enuc"])


@pytest.mark.parametrize("R", list(GOLDEN_CCSD))
def test_full_ci_of_the_oracle_integrals_reproduces_the_notebook_ccsd_energies(R):
    _, I = _h2(R)
    assert abs(_full_ci(I) - GOLDEN_CCSD[R]) < 3e-7


def test_full_eri_tensor_against_the_notebook_ccsd_energy_and_density_matrix():
    """For two electrons CCSD is exact, so a 16 x 16 full-CI of the oracle integrals must give the notebook's
    E(CCSD) = -1.151672678339737 and its printed AO density matrix (both carry pyscf's CCSD convergence
    tolerance of 1e-7): this pins every element class of the ERI tensor, not only its contraction with one dm."""
    _, I = _h2(0.74)
    _, C = scf_ref.generalized_eigh(I["h1e"], I["s1e"])
    h = C.T @ I["h1e"] @ C
    e = np.einsum("pi,qj,rk,sl,pqrs->ijkl", C, C, C, C, I["eri"], optimize=True)
    n = 4
    eye = np.eye(n)
    H = (np.einsum("ik,jl->ijkl", h, eye) + np.einsum("ik,jl->ijkl", eye, h) + e.transpose(0, 2, 1, 3)).reshape(n * n, n * n)
    ev, vec = np.linalg.eigh(H)
    c = vec[:, 0].reshape(n, n)
    assert np.abs(c - c.T).max() < 1e-12  # singlet ground state
    assert abs(ev[0] + I["enuc"] - (-1.151672678339737)) < 3e-7
    dm = C @ (2.0 * c @ c.T) @ C.T
    printed = np.array([[0.23211218, 0.18285689, 0.20607785, 0.16169728], [0.18285689, 0.1546802, 0.16169728, 0.133479],
                        [0.20607785, 0.16169728, 0.23211218, 0.18285689], [0.16169728, 0.133479, 0.18285689, 0.1546802]])
    assert np.abs(dm - printed).max() < 1e-6


def test_oracle_eigensolver_and_degenerate_cotangent():
    rng = np.random.default_rng(0)
    n = 6
    A = rng.standard_normal((n, n))
    A = A + A.T
    Bm = rng.standard_normal((n, n))
    Bm = Bm @ Bm.T + n * np.eye(n)
    w, V = scf_ref.generalized_eigh(A, Bm)
    assert np.allclose(A @ V, Bm @ V * w, atol=1e-10) and np.allclose(V.T @ Bm @ V, np.eye(n), atol=1e-10)
    # cotangent of a function of the eigenvalues only: d sum(w^2)/dA = 2 V diag(w) V^T = 2A (B = I)
    w, V = np.linalg.eigh(A)
    assert np.allclose(scf_ref.degen_eigh_bwd(w, V, 2 * w, None), 2 * A, atol=1e-10)
    # exactly degenerate pair: the 1/(l_j - l_i) term is dropped instead of producing inf/nan
    D = np.diag([1.0, 1.0, 3.0])
    w, V = np.linalg.eigh(D)
    g = scf_ref.degen_eigh_bwd(w, V, None, rng.standard_normal((3, 3)))
    assert np.isfinite(g).all()


def test_oracle_diis_matches_plain_iteration_at_convergence():
    _, I = _h2(0.9)
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    e_diis, _, _ = scf_ref.rhf_loop(dm0, I["eri"], I["s1e"], I["h1e"], I["enuc"], 2, max_cycle=30)
    e_plain, _, _ = scf_ref.rhf_loop(dm0, I["eri"], I["s1e"], I["h1e"], I["enuc"], 2, max_cycle=60, diis_start_cycle=10**6)
    assert abs(e_diis - e_plain) < 1e-10


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("R", list(GOLDEN_RHF))
def test_cuda_rhf_reproduces_the_reference_notebook_energies(R):
    """J and K from csrc/jk.cu inside the torch SCF loop against numbers the reference itself printed."""
    import torch

    from qex_b200 import scf

    _, I = _h2(R)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    dm0 = scf.core_guess(t["h1e"], t["s1e"], 2)
    e, dm, hist = scf.rhf_loop(dm0, t["eri"], t["s1e"], t["h1e"], I["enuc"], 2, max_cycle=30)
    assert abs(e.item() - GOLDEN_RHF[R]) < 1e-10
    _, dm_ref, hist_ref = scf_ref.rhf_loop(scf_ref.core_guess(I["h1e"], I["s1e"], 2), I["eri"], I["s1e"], I["h1e"],
                                           I["enuc"], 2, max_cycle=30)
    assert np.abs(hist.cpu().numpy() - hist_ref).max() < 1e-10
    assert np.abs(dm.cpu().numpy() - dm_ref).max() < 1e-9


def _ks_problem(R=0.74):
    m, I = _h2(R)
    g = gen_grid.Grids(m, n_rad=31, n_theta=5, n_phi=4).build()
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    return m, I, g, spec, theta


@pytest.mark.gpu
def test_cuda_ks_scf_loop_matches_oracle_loop():
    """15-cycle KS loop with the LocalMLP functional: every cycle's total energy against the numpy loop."""
    import torch

    from qex_b200 import _lib, scf
    from qex_b200.engine import NetSpec, XCContext

    m, I, g, spec, theta = _ks_problem()
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)

    def exc_vrho(rho):
        return mlp_ref.exc_and_vrho_local(spec, theta, rho)

    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    xc = XCContext(nao=4, ngrids_max=g.size, ncomp=1, net=NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64))
    xc.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    th = torch.as_tensor(theta).cuda()
    args = (I["eri"], ao, g.weights, I["s1e"], I["h1e"], I["enuc"], 2, exc_vrho)
    targs = (t["eri"], t["s1e"], t["h1e"], I["enuc"], 2)
    # (a) plain fixed-point iteration (DIIS switched off): strict parity on all 15 cycles
    _, dm_ref, hist_ref = scf_ref.scf_loop(dm0, *args, diis_start_cycle=10**6)
    _, dm, hist = scf.scf_loop(xc, th, torch.as_tensor(dm0).cuda(), *targs, diis_start_cycle=10**6)
    assert np.abs(hist.detach().cpu().numpy() - hist_ref).max() < 1e-10
    assert np.abs(dm.detach().cpu().numpy() - dm_ref).max() < 1e-9
    # (b) with DIIS (reference defaults): identical until the extrapolation starts; after that the DIIS
    # system is ill-conditioned (nearly dependent error vectors) and LAPACK vs cuSOLVER rounding is
    # amplified, in the reference as much as here -- the energies agree to ~1e-8, not to 1e-10
    _, dm_ref, hist_ref = scf_ref.scf_loop(dm0, *args)
    _, dm, hist = scf.scf_loop(xc, th, torch.as_tensor(dm0).cuda(), *targs)
    d = np.abs(hist.detach().cpu().numpy() - hist_ref)
    assert d[:3].max() < 1e-12 and d.max() < 1e-6


@pytest.mark.gpu
def test_cuda_scf_gradient_wrt_theta_matches_finite_differences():
    """d e_tot / d theta through 4 SCF cycles (XC VJP + J reverse + degenerate-safe eigh rule) against
    central differences of the numpy oracle loop along two random directions.  Strict with the plain
    iteration; with DIIS the extrapolation solve is ill-conditioned, which makes the finite difference
    itself noisy (the same 1e-5 scatter shows up with a pure-torch CPU loop), so that leg is loose."""
    import torch

    from qex_b200 import _lib, scf
    from qex_b200.engine import NetSpec, XCContext

    m, I, g, spec, theta = _ks_problem(0.9)
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)

    def e_of(th, **kw):
        def exc_vrho(rho):
            return mlp_ref.exc_and_vrho_local(spec, th, rho)
        return scf_ref.scf_loop(dm0, I["eri"], ao, g.weights, I["s1e"], I["h1e"], I["enuc"], 2, exc_vrho, **kw)[0]

    xc = XCContext(nao=4, ngrids_max=g.size, ncomp=1, net=NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64))
    xc.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    th = torch.as_tensor(theta).cuda().requires_grad_(True)
    rng = np.random.default_rng(5)
    for kw, tol in ((dict(max_cycle=4, diis_start_cycle=10**6), 1e-7), (dict(max_cycle=4), 1e-3)):
        e, _, _ = scf.scf_loop(xc, th, torch.as_tensor(dm0).cuda(), t["eri"], t["s1e"], t["h1e"], I["enuc"], 2, **kw)
        (grad,) = torch.autograd.grad(e, th)
        grad = grad.cpu().numpy()
        for _ in range(2):
            d = rng.standard_normal(theta.shape)
            d /= np.linalg.norm(d)
            h = 1e-5
            fd = (e_of(theta + h * d, **kw) - e_of(theta - h * d, **kw)) / (2 * h)
            assert abs(fd - grad @ d) < tol * max(1.0, abs(fd))


@pytest.mark.gpu
def test_cuda_implicit_differentiation_of_the_scf_fixed_point():
    """`make_implicit_diff(_scf, optimality_cond=_scf_optimality_cond, solver=gen_gmres())` (hf_legacy.py:202-205):
    gradient of a density-matrix loss at self-consistency from ONE adjoint GMRES solve whose operator is the VJP
    of the XC + J kernels, against central differences of the oracle's converged fixed point."""
    import torch

    from qex_b200 import _lib, scf
    from qex_b200.engine import NetSpec, XCContext

    m, I, g, spec, theta = _ks_problem(0.9)
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    rng = np.random.default_rng(1)
    M = rng.standard_normal((4, 4))
    M = M + M.T

    def loss_ref(th):
        dm, delta = scf_ref.scf_fixed_point(dm0, I["eri"], ao, g.weights, I["s1e"], I["h1e"], 2,
                                            lambda rho: mlp_ref.exc_and_vrho_local(spec, th, rho))
        assert delta < 1e-10
        return float((dm * M).sum())

    xc = XCContext(nao=4, ngrids_max=g.size, ncomp=1, net=NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64))
    xc.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    th = torch.as_tensor(theta).cuda().requires_grad_(True)
    dm = scf.scf_fixed_point(xc, th, torch.as_tensor(dm0).cuda(), t["eri"], t["s1e"], t["h1e"], 2)
    loss = (dm * torch.as_tensor(M).cuda()).sum()
    assert abs(loss.item() - loss_ref(theta)) < 1e-9
    (grad,) = torch.autograd.grad(loss, th)
    grad = grad.cpu().numpy()
    for _ in range(2):
        d = rng.standard_normal(theta.shape)
        d /= np.linalg.norm(d)
        h = 1e-4
        fd = (loss_ref(theta + h * d) - loss_ref(theta - h * d)) / (2 * h)
        assert abs(fd - grad @ d) < 1e-8 + 1e-5 * abs(fd)


def test_host_integral_generator_matches_the_oracle_integrals():
    from qex_b200 import ints

    m, I = _h2(1.2)
    J = ints.integrals(m._atm, m._bas, m._env)
    for k in ("s1e", "t1e", "v1e", "h1e", "eri"):
        assert np.abs(J[k] - I[k]).max() < 1e-14
    assert abs(J["enuc"] - I["enuc"]) < 1e-15


@pytest.mark.gpu
def test_cuda_batched_scf_matches_per_molecule_loops_and_oracle():
    """Config c4 in miniature: three H2 geometries in ONE batched loop (batched XC launches, batched J kernel,
    batched eigensolver) against three single-molecule GPU loops and against the numpy oracle loop."""
    import torch

    from qex_b200 import _lib, hf, scf
    from qex_b200.engine import NetSpec, XCContext

    bonds = [0.74, 0.5, 1.5]
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    mols, Is, grids = [], [], []
    for R in bonds:
        m, I = _h2(R)
        mols.append(m)
        Is.append(I)
        grids.append(gen_grid.Grids(m, n_rad=31, n_theta=5, n_phi=4).build())
    G = grids[0].size
    net = NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64)
    xb = XCContext(nao=4, ngrids_max=G, ncomp=1, nbatch=3, net=net)
    xb.set_grid(np.stack([g.coords for g in grids]), np.stack([g.weights for g in grids]))
    xb.set_basis(mols[0]._atm, mols[0]._bas, np.stack([m._env for m in mols])).eval_ao(0)
    st = lambda k: torch.as_tensor(np.stack([I[k] for I in Is])).cuda()  # noqa: E731
    dm0 = np.stack([scf_ref.core_guess(I["h1e"], I["s1e"], 2) for I in Is])
    enuc = torch.as_tensor(np.array([I["enuc"] for I in Is])).cuda()
    th = torch.as_tensor(theta).cuda()
    # batched J/K kernel against the oracle einsums
    vj, vk = hf.dot_eri_dm_batched(st("eri"), torch.as_tensor(dm0).cuda())
    for b, I in enumerate(Is):
        rj, rk = scf_ref.jk_ref.dot_eri_dm(I["eri"], dm0[b])
        assert np.abs(vj[b].cpu().numpy() - rj).max() < 1e-12 and np.abs(vk[b].cpu().numpy() - rk).max() < 1e-12
    kw = dict(max_cycle=8, diis_start_cycle=10**6)
    e_b, dm_b, hist_b = scf.scf_loop_batched(xb, th, torch.as_tensor(dm0).cuda(), st("eri"), st("s1e"), st("h1e"), enuc, 2, **kw)
    for b, (m, I, g) in enumerate(zip(mols, Is, grids)):
        x1 = XCContext(nao=4, ngrids_max=G, ncomp=1, net=net)
        x1.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
        t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
        _, dm1, hist1 = scf.scf_loop(x1, th, torch.as_tensor(dm0[b]).cuda(), t["eri"], t["s1e"], t["h1e"], I["enuc"], 2, **kw)
        assert (hist_b[:, b] - hist1).abs().max().item() < 1e-11
        ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
        _, dm_ref, hist_ref = scf_ref.scf_loop(dm0[b], I["eri"], ao, g.weights, I["s1e"], I["h1e"], I["enuc"], 2,
                                               lambda rho: mlp_ref.exc_and_vrho_local(spec, theta, rho), **kw)
        assert np.abs(hist_b[:, b].cpu().numpy() - hist_ref).max() < 1e-10
        assert np.abs(dm_b[b].cpu().numpy() - dm_ref).max() < 1e-9
    # with DIIS (reference defaults): same statement as the single-molecule test
    _, _, hist_d = scf.scf_loop_batched(xb, th, torch.as_tensor(dm0).cuda(), st("eri"), st("s1e"), st("h1e"), enuc, 2)
    assert torch.isfinite(hist_d).all() and hist_d.shape == (15, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nb", [(1, 3), (2, 5), (4, 64), (7, 9), (8, 33), (13, 4), (16, 70)])
def test_cuda_small_generalized_eigh_kernel(n, nb):
    """csrc/eigh.cu (one thread per matrix, Jacobi) against the oracle's LAPACK route: eigenvalues, the defining
    relations A V = B V diag(w), V^T B V = I, and the cotangent of a sign-invariant function against torch.linalg."""
    import torch

    from qex_b200 import scf

    rng = np.random.default_rng(n * 100 + nb)
    A = rng.standard_normal((nb, n, n))
    A = A + A.transpose(0, 2, 1)
    Bm = rng.standard_normal((nb, n, n))
    Bm = Bm @ Bm.transpose(0, 2, 1) + n * np.eye(n)
    if n >= 2:
        A[0] = np.diag(np.arange(n) // 2).astype(float)  # exactly degenerate pairs
        Bm[0] = np.eye(n)
    At, Bt = torch.as_tensor(A).cuda().requires_grad_(True), torch.as_tensor(Bm).cuda()
    w, V = scf.generalized_eigh_batched(At, Bt, small_kernel=True)
    wn, Vn = w.detach().cpu().numpy(), V.detach().cpu().numpy()
    for b in range(nb):
        w_ref, _ = scf_ref.generalized_eigh(A[b], Bm[b])
        assert np.abs(wn[b] - w_ref).max() <= 1e-12 * max(1.0, np.abs(w_ref).max())
        assert np.abs(A[b] @ Vn[b] - Bm[b] @ Vn[b] * wn[b]).max() <= 1e-11 * max(1.0, np.abs(A[b]).max())
        assert np.abs(Vn[b].T @ Bm[b] @ Vn[b] - np.eye(n)).max() <= 1e-12
    # gradient of a function of the occupied projector and the eigenvalues (invariant to signs / rotations in
    # non-degenerate subspaces): kernel + custom rule vs torch.linalg route + autograd
    nocc = max(1, n // 2)
    M = torch.as_tensor(rng.standard_normal((n, n))).cuda()
    c = torch.as_tensor(rng.standard_normal(n)).cuda()

    def f(w_, V_):
        P = V_[..., :nocc] @ V_[..., :nocc].transpose(-1, -2)
        return (P * (M + M.T)).sum() + (w_ * c).sum()

    (g1,) = torch.autograd.grad(f(w[1:], V[1:]), At)
    A2 = torch.as_tensor(A).cuda().requires_grad_(True)
    w2, V2 = scf.generalized_eigh_batched(A2, Bt, small_kernel=False)
    (g2,) = torch.autograd.grad(f(w2[1:], V2[1:]), A2)
    assert (g1 - g2).abs().max().item() <= 1e-9 * max(1.0, g2.abs().max().item())
