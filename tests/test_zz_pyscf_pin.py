"""Pins of the XC grid-integration path against numbers the REFERENCE ITSELF holds.

notebooks/04_notebook_td_trainer.ipynb (cell 1 / cell 2 outputs) freezes, for H2 / 6-31G at 0.74 A on the pyscf
grid the trainers use (level 0, Stratmann partition: trainer_legacy_no_jit.py:248-251, dataset_generation.py:139-142):

* "converged SCF energy = -1.03718794786902" of an RKS run with ``xc = "lda"`` (dataset_generation.py:385-389), on a
  grid of 1240 points ("Number of grid points: 1240");
* an older run's 1192-point grid with the first/last three densities of the CCSD density matrix it prints.

Everything between the molecule and those numbers is on the path: grid generation, AO values, rho, E_xc, V_xc
assembly, the SCF loop.  The oracle reproduces the energy to < 1e-12 Ha (CPU tests below), the CUDA kernels
(AO evaluator, rho contraction, V_xc assembly through the reference-facing ``NumInt.nr_rks``) to < 1e-9 Ha.
(The file name sorts last on purpose.)"""
import numpy as np
import pytest

from oracle import grid_ref, gto_ref, ints_ref, scf_ref
from qex_b200 import gen_grid, gto

E_LDA_NOTEBOOK = -1.03718794786902
DM_CCSD_NOTEBOOK = np.array([[0.23211218, 0.18285689, 0.20607785, 0.16169728],
                             [0.18285689, 0.1546802, 0.16169728, 0.133479],
                             [0.20607785, 0.16169728, 0.23211218, 0.18285689],
                             [0.16169728, 0.133479, 0.18285689, 0.1546802]])
RHO_TAIL_NOTEBOOK = (6.72815113e-13, 3.23056307e-11, 1.63724571e-12)


def _h2():
    m = gto.h2(0.74, "6-31g")
    return m, ints_ref.integrals(m._atm, m._bas, m._env)


def _nearest_rel(values, target):
    return float(np.abs(values / target - 1.0).min())


def test_oracle_grid_has_the_point_counts_the_notebook_logs():
    m, _ = _h2()
    c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0)
    assert c.shape == (1240, 3) and w.shape == (1240,)
    c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, xi_table=None)
    assert c.shape == (1192, 3)
    # single atom: the partition is the identity and the weights integrate a Gaussian (level 3: 50 x 302 pruned)
    c, w = grid_ref.build([1], [[0.0, 0.0, 0.0]], level=3)
    assert abs((np.exp(-(c**2).sum(1)) * w).sum() - np.pi**1.5) < 1e-9


def test_oracle_lda_rks_energy_matches_the_notebook():
    m, I = _h2()
    c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, becke_scheme=grid_ref.stratmann)
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, c, 0)
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    e, dm, hist = scf_ref.scf_loop(dm0, I["eri"], ao, w, I["s1e"], I["h1e"], I["enuc"], 2, grid_ref.lda_exchange,
                                   max_cycle=30)
    assert abs(hist[-1] - hist[-2]) < 1e-13
    assert abs(e - E_LDA_NOTEBOOK) < 1e-12
    # the plain-Becke partition or the xi = 1 radial map miss the printed value by 3e-4 / 3e-3 Ha: the pin is sharp
    c2, w2 = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, becke_scheme=grid_ref.original_becke)
    ao2 = gto_ref.eval_ao(m._atm, m._bas, m._env, c2, 0)
    e2, _, _ = scf_ref.scf_loop(dm0, I["eri"], ao2, w2, I["s1e"], I["h1e"], I["enuc"], 2, grid_ref.lda_exchange, max_cycle=30)
    assert abs(e2 - E_LDA_NOTEBOOK) > 1e-4


def test_oracle_ao_and_rho_reproduce_the_notebook_tail_densities():
    m, _ = _h2()
    c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, xi_table=None)  # the 1192-point run
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, c, 0)
    rho = np.einsum("gi,ij,gj->g", ao, DM_CCSD_NOTEBOOK, ao)
    for t in RHO_TAIL_NOTEBOOK:
        assert _nearest_rel(rho, t) < 3e-8  # 8 printed digits of dm, 9 of rho


def test_box_ordering_reproduces_the_head_and_tail_the_notebook_prints():
    """pyscf hands the points out sorted into spatial boxes; cell 2 of the notebook prints `Density: [a b c ... c b a]`.
    With the box sort restated, rho[:3] and rho[-3:] come out as printed, in order (oracle and host generator)."""
    m, _ = _h2()
    c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, xi_table=None, sort_grids=True)
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, c, 0)
    rho = np.einsum("gi,ij,gj->g", ao, DM_CCSD_NOTEBOOK, ao)
    printed = np.array(RHO_TAIL_NOTEBOOK)
    assert np.abs(rho[:3] / printed - 1).max() < 3e-8 and np.abs(rho[-3:] / printed[::-1] - 1).max() < 3e-8
    # the unsorted grid starts at the nucleus instead
    c0, _ = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, xi_table=None)
    assert np.linalg.norm(c0[0]) < 0.01 and np.linalg.norm(c[0]) > 8.0
    # host generator: same permutation (independent implementation of the box key), same weights
    g = gen_grid.Grids(m)
    g.level = 0
    g.becke_scheme = gen_grid.stratmann
    g.build(sort_grids=True)
    c1, w1 = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, sort_grids=True)
    assert np.abs(g.coords - c1).max() < 1e-13 and np.abs(g.weights - w1).max() < 1e-13
    assert abs(g.weights.sum() - g.build().weights.sum()) < 1e-12


def test_host_grid_generator_matches_the_oracle_grid():
    m, _ = _h2()
    for kw in ({}, {"becke_scheme": "becke"}, {"level": 1}):
        g = gen_grid.Grids(m)
        g.level = kw.get("level", 0)
        g.becke_scheme = gen_grid.original_becke if "becke_scheme" in kw else gen_grid.stratmann
        g.build()
        c, w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=g.level,
                              becke_scheme=grid_ref.original_becke if "becke_scheme" in kw else grid_ref.stratmann)
        assert g.coords.shape == c.shape
        assert np.abs(g.coords - c).max() < 1e-13 and np.abs(g.weights - w).max() < 1e-13
    # heteronuclear: the atomic-size adjustment enters
    mol = gto.Mole([("O", (0.0, 0.0, 0.0)), ("H", (0.0, 0.757, 0.587)), ("H", (0.0, -0.757, 0.587))],
                   basis=gto.even_tempered_basis([2, 1]))
    g = gen_grid.Grids(mol)
    g.level = 0
    g.build()
    c, w = grid_ref.build(mol.atom_charges(), mol.atom_coords(), level=0, becke_scheme=grid_ref.original_becke)
    assert np.abs(g.coords - c).max() < 1e-13 and np.abs(g.weights - w).max() < 1e-13
    assert abs((np.exp(-((g.coords - mol.atom_coords()[0]) ** 2).sum(1)) * g.weights).sum() - np.pi**1.5) < 1e-2
    # pyscf's calling conventions: build(mol) positionally, atom_grid keyed by element symbol, prune switched off
    g2 = gen_grid.Grids(None)
    g2.atom_grid = {"H": (12, 26), "O": (15, 50)}
    g2.prune = None
    g2.build(mol, with_non0tab=True)
    assert g2.size == 2 * 12 * 26 + 15 * 50
    assert abs((np.exp(-((g2.coords - mol.atom_coords()[1]) ** 2).sum(1)) * g2.weights).sum() - np.pi**1.5) < 5e-2


# ---------------------------------------------------------------------------------------------- CUDA path
def _slater_eval_xc(xc_code, rho, *args, **kwargs):
    """libxc LDA_X in the return structure of ``eval_xc`` (exc per particle, vrho = d(rho exc)/d rho)."""
    exc, vrho = grid_ref.lda_exchange(rho)
    return exc, (vrho, None, None, None), None, None


@pytest.mark.gpu
def test_cuda_lda_rks_energy_matches_the_notebook(lib):
    """The reference's own data-generation step -- ``dft.RKS(mol); mf.grids = level-0 Stratmann grid; mf.xc = "lda";
    mf.kernel()`` (dataset_generation.py:377-389) -- with the XC part on the CUDA kernels through the reference-facing
    ``NumInt.nr_rks`` (AO evaluator K1, rho contraction K2, V_xc assembly K5) and J on the J/K kernel; the host loop is
    the oracle's DIIS / eigensolver.  Must land on the energy the reference's notebook printed."""
    import torch

    from qex_b200 import hf
    from qex_b200.numint import NumInt

    m, I = _h2()
    g = gen_grid.Grids(m)
    g.level = 0
    g.becke_scheme = gen_grid.stratmann
    g.build(device=0)
    assert g.size == 1240
    ni = NumInt(cache_ao=True)
    ni.eval_xc = _slater_eval_xc
    eri = torch.tensor(I["eri"], device="cuda")

    def veff(dm):
        nelec, exc, vxc = ni.nr_rks(m, g, "LDA", dm, hermi=1)
        vj, _ = hf.dot_eri_dm(eri, torch.tensor(dm, device="cuda"), with_j=True, with_k=False)
        return vj.cpu().numpy() + vxc, exc, vj.cpu().numpy(), nelec

    dm = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    vhf, exc, vj, _ = veff(dm)
    st = scf_ref.initialize_diis(15)
    e_hist = []
    for cycle in range(25):
        fock = I["h1e"] + vhf
        if cycle >= 1:
            fock, st = scf_ref.apply_diis(st, fock, dm, I["s1e"], 15, 2, 0.0)
        mo_e, mo_c = scf_ref.generalized_eigh(fock, I["s1e"])
        dm = scf_ref.make_rdm1(mo_c, scf_ref.get_occ(2, mo_e))
        vhf, exc, vj, nelec = veff(dm)
        e_hist.append(scf_ref.energy_tot(dm, I["h1e"], vj, exc, I["enuc"]))
    assert abs(e_hist[-1] - e_hist[-2]) < 1e-12
    assert abs(e_hist[-1] - E_LDA_NOTEBOOK) < 1e-9        # north_star: total SCF energy within 1e-8 Ha
    assert abs(nelec - 2.0) < 5e-3                         # level-0 grid: N_elec is only this good in pyscf too


@pytest.mark.gpu
def test_cuda_ao_and_rho_reproduce_the_notebook_tail_densities(lib):
    """``numint.eval_ao(mol, coords)`` + ``numint.eval_rho(mol, ao, dm_ao)`` (dataset_generation.py:392-395) on the CUDA
    kernels, on the 1192-point grid of the notebook's older cells, against the densities it prints."""
    from qex_b200 import numint

    m, _ = _h2()
    c, _w = grid_ref.build(m.atom_charges(), m.atom_coords(), level=0, xi_table=None)
    ao = numint.eval_ao(m, c, deriv=0)
    rho = numint.eval_rho(m, ao, DM_CCSD_NOTEBOOK, xctype="LDA")
    assert rho.shape == (1192,)
    for t in RHO_TAIL_NOTEBOOK:
        assert _nearest_rel(rho, t) < 3e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["h2_stratmann", "h2o_becke_adjust", "c120_stratmann", "c120_becke"])
def test_cuda_becke_partition_matches_the_host_partition(lib, case):
    """qexxc_becke_partition (csrc/grid.cu) against the NumPy partition and the oracle grid; 120 atoms exceed the
    shared-memory inverse-distance table, so both kernel variants run."""
    if case.startswith("h2_"):
        mol = gto.h2(0.74, "6-31g")
    elif case.startswith("h2o"):
        mol = gto.Mole([("O", (0.0, 0.0, 0.0)), ("H", (0.0, 0.757, 0.587)), ("H", (0.0, -0.757, 0.587))],
                       basis=gto.even_tempered_basis([2, 1]))
    else:
        mol = gto.synthetic_molecule(120, [1], seed=3)
    scheme = gen_grid.stratmann if "stratmann" in case else gen_grid.original_becke
    grids = []
    for device in (None, 0):
        g = gen_grid.Grids(mol)
        g.level = 0
        g.becke_scheme = scheme
        if case.startswith("c120"):
            g.atom_grid = {6: (6, 26)}  # 156 points per atom keep the NumPy partition (G x 120^2 pair terms) short
        grids.append(g.build(device=device))
    host, dev = grids
    assert np.array_equal(host.coords, dev.coords)
    scale = np.abs(host.weights).max()
    assert np.abs(dev.weights - host.weights).max() <= 1e-13 * scale
    if not case.startswith("c120"):
        c, w = grid_ref.build(mol.atom_charges(), mol.atom_coords(), level=0,
                              becke_scheme=grid_ref.stratmann if "stratmann" in case else grid_ref.original_becke)
        assert np.abs(dev.weights - w).max() <= 1e-13 * scale
    assert lib.qexxc_grid_launch_count() > 0


def test_lda_code_handling_on_the_host():
    from qex_b200 import numint, xc

    assert numint._xctype(numint.NumInt(), "lda") == "LDA" and numint._xctype(numint.NumInt(), "LDA,") == "LDA"
    with pytest.raises(NotImplementedError):
        numint._xctype(numint.NumInt(), "b3lyp")
    with pytest.raises(NotImplementedError):
        xc.lda_eval_xc("lda,vwn", np.ones(4))
    with pytest.raises(NotImplementedError):
        xc.lda_eval_xc("lda", np.ones(4), spin=1)


@pytest.mark.gpu
def test_cuda_lda_exchange_kernel_matches_the_closed_form(lib):
    import torch

    from qex_b200 import xc

    rng = np.random.default_rng(0)
    rho = np.concatenate([10.0 ** rng.uniform(-14, 1, 5001), [0.0, -1e-18, 1e-300]])  # 5004 = 3 x 1668
    e_ref, v_ref = grid_ref.lda_exchange(rho)
    e, (v, a, b, c), f, k = xc.lda_eval_xc("lda", rho)
    assert (a, b, c, f, k) == (None,) * 5 and isinstance(e, np.ndarray)
    assert np.abs(e - e_ref).max() <= 4e-16 * np.abs(e_ref).max() and np.abs(v - v_ref).max() <= 4e-16 * np.abs(v_ref).max()
    et, vt = xc.lda_exchange(torch.tensor(rho, device="cuda").reshape(3, -1))  # tensors stay on the device
    assert et.is_cuda and et.shape == (3, rho.size // 3) and np.array_equal(et.cpu().numpy().ravel(), e)
    assert lib.qexxc_lda_launch_count() >= 2


@pytest.mark.gpu
def test_cuda_device_resident_lda_scf_matches_the_notebook(lib):
    """The same pin with NOTHING on the host: `scf.scf_loop` (J kernel, batched-Jacobi / cuSOLVER eigensolver, DIIS) around
    rho (K2) -> Slater exchange (csrc/xc_lda.cu) -> V_xc assembly (K5), grid from the CUDA partition.  Then the built-in
    `nr_rks(mol, grids, "lda", dm)` against the host-callback route, and three bond lengths in ONE batched loop."""
    import torch

    from qex_b200 import scf
    from qex_b200.engine import XCContext
    from qex_b200.numint import NumInt

    def problem(R):
        m = gto.h2(R, "6-31g")
        g = gen_grid.Grids(m)
        g.level = 0
        g.becke_scheme = gen_grid.stratmann
        return m, ints_ref.integrals(m._atm, m._bas, m._env), g.build(device=0)

    m, I, g = problem(0.74)
    ctx = XCContext(nao=4, ngrids_max=g.size, ncomp=1)
    ctx.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    dm0 = torch.as_tensor(scf_ref.core_guess(I["h1e"], I["s1e"], 2)).cuda()
    e, dm, hist = scf.scf_loop(ctx, None, dm0, t["eri"], t["s1e"], t["h1e"], I["enuc"], 2, xctype="LDA", max_cycle=25)
    assert abs(float(hist[-1] - hist[-2])) < 1e-11
    assert abs(float(e) - E_LDA_NOTEBOOK) < 1e-9
    # built-in functional of nr_rks == the host-callback route of the first pin test
    ni, nj = NumInt(), NumInt()
    nj.eval_xc = _slater_eval_xc
    d = dm.cpu().numpy()
    n1, e1, v1 = ni.nr_rks(m, g, "lda", d, hermi=1)
    n2, e2, v2 = nj.nr_rks(m, g, "LDA", d, hermi=1)
    assert abs(n1 - n2) < 1e-13 and abs(e1 - e2) < 1e-13 and np.abs(v1 - v2).max() < 1e-13
    # batched: three geometries in one loop against the oracle loop on the oracle grid
    bonds = [0.74, 0.5, 1.5]
    P = [problem(R) for R in bonds]
    xb = XCContext(nao=4, ngrids_max=P[0][2].size, ncomp=1, nbatch=3)
    xb.set_grid(np.stack([p[2].coords for p in P]), np.stack([p[2].weights for p in P]))
    xb.set_basis(P[0][0]._atm, P[0][0]._bas, np.stack([p[0]._env for p in P])).eval_ao(0)
    st = lambda k: torch.as_tensor(np.stack([p[1][k] for p in P])).cuda()  # noqa: E731
    dmb = torch.as_tensor(np.stack([scf_ref.core_guess(p[1]["h1e"], p[1]["s1e"], 2) for p in P])).cuda()
    enuc = torch.as_tensor(np.array([p[1]["enuc"] for p in P])).cuda()
    eb, _, _ = scf.scf_loop_batched(xb, None, dmb, st("eri"), st("s1e"), st("h1e"), enuc, 2, xctype="LDA", max_cycle=25)
    for b, (mm, II, gg) in enumerate(P):
        ao = gto_ref.eval_ao(mm._atm, mm._bas, mm._env, gg.coords, 0)
        e_ref, _, _ = scf_ref.scf_loop(scf_ref.core_guess(II["h1e"], II["s1e"], 2), II["eri"], ao, gg.weights, II["s1e"],
                                       II["h1e"], II["enuc"], 2, grid_ref.lda_exchange, max_cycle=25)
        assert abs(float(eb[b]) - e_ref) < 1e-9
    assert abs(float(eb[0]) - E_LDA_NOTEBOOK) < 1e-9


@pytest.mark.gpu
def test_nr_rks_lda_like_the_reference_pyscf_consistency_test(lib):
    """tests/test_numint.py:208-258 of the reference (`test_pyscf_pyscfad_consistency`): H2 at 1.0 A, STO-3G, xc "LDA",
    grids.level = 1 (pyscf's default Becke scheme), `ni.nr_rks(mol, grids, "LDA", dm)` compared between two
    implementations at rtol 1e-6 / 1e-5.  Here: the CUDA path against the oracle (pyscf itself is absent), same calls,
    tolerances five orders tighter; then the converged RKS energies of both."""
    import torch

    from qex_b200 import scf
    from qex_b200.engine import XCContext
    from qex_b200.numint import NumInt
    from oracle import numint_ref

    mol = gto.h2(1.0, "sto-3g")
    grids = gen_grid.Grids(mol)
    grids.level = 1
    grids.build(device=0)
    c, w = grid_ref.build(mol.atom_charges(), mol.atom_coords(), level=1, becke_scheme=grid_ref.original_becke)
    assert grids.coords.shape == c.shape and np.abs(grids.weights - w).max() < 1e-13 * np.abs(w).max()
    I = ints_ref.integrals(mol._atm, mol._bas, mol._env)
    dm = scf_ref.core_guess(I["h1e"], I["s1e"], 2)  # stands in for get_init_guess()
    ni = NumInt()
    nelec, exc, vxc = ni.nr_rks(mol, grids, "LDA", dm)
    ao = gto_ref.eval_ao(mol._atm, mol._bas, mol._env, c, 0)
    n0, e0, v0 = numint_ref.nr_rks(ao, w, dm, _slater_eval_xc, "NN")  # the LDA branch shares the NN branch's assembly
    np.testing.assert_allclose(nelec, n0, rtol=1e-11)
    np.testing.assert_allclose(exc, e0, rtol=1e-11)
    np.testing.assert_allclose(vxc, v0, rtol=1e-10, atol=1e-13)
    assert abs(nelec - 2.0) < 1e-4  # a level-1 grid integrates the density to 1e-5
    ctx = XCContext(nao=2, ngrids_max=grids.size, ncomp=1)
    ctx.set_grid(grids.coords, grids.weights).set_basis(mol._atm, mol._bas, mol._env).eval_ao(0)
    t = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in I.items() if k != "enuc"}
    e_gpu, _, _ = scf.scf_loop(ctx, None, torch.as_tensor(dm).cuda(), t["eri"], t["s1e"], t["h1e"], I["enuc"], 2,
                               xctype="LDA", max_cycle=20)
    e_ref, _, _ = scf_ref.scf_loop(dm, I["eri"], ao, w, I["s1e"], I["h1e"], I["enuc"], 2, grid_ref.lda_exchange, max_cycle=20)
    np.testing.assert_allclose(float(e_gpu), e_ref, rtol=1e-10)


@pytest.mark.parametrize("z", [1, 8])
def test_level_tables_give_converging_atomic_quadratures(z):
    """Rows of the level / xi / pruning tables that no reference number pins (only H at level 0 is): every level must at
    least be a quadrature that converges -- a Gaussian, a 1s Slater density and an l = 4 moment on one atom."""
    errs = []
    for level in (0, 1, 3, 5):
        c, w = grid_ref.build([z], [[0.0, 0.0, 0.0]], level=level)
        r2 = (c**2).sum(1)
        errs.append((abs((np.exp(-r2) * w).sum() - np.pi**1.5),
                     abs((np.exp(-2.0 * np.sqrt(r2)) / np.pi * w).sum() - 1.0),
                     abs((c[:, 0] ** 2 * c[:, 2] ** 2 * np.exp(-r2) * w).sum() - np.pi**1.5 / 4.0)))
        g = gen_grid.Grids(gto.Mole([(z, (0.0, 0.0, 0.0))], basis=gto.even_tempered_basis([1]), unit="Bohr"))
        g.level = level
        g.build()
        assert g.size == w.size and np.abs(g.weights - w).max() < 1e-13 * np.abs(w).max()
    errs = np.array(errs)
    assert (errs[0] < 1e-2).all() and (errs[1] < 1e-8).all() and (errs[2] < 1e-9).all() and (errs[3] < 1e-10).all()
