"""GPU parity of the INT8 (exact digit-split, tcgen05) contractions -- `csrc/contract_i8.cu` -- through the C ABI:
against the float64 CPU oracle at small sizes, against the FP64 DMMA path at medium sizes, and on inputs with a wide
dynamic range (the block-floating-point exponents are what these stress).

Same bar as the FP64 path (BASELINE.json north_star): V_xc / gradient elements within 1e-10 relative to the largest
element; the path is the default for nao >= 256 and forced here with QEXXC_I8=1.
"""
import numpy as np
import pytest

from oracle import numint_ref
from tests._util import rel_err, synth_problem

pytestmark = pytest.mark.gpu

TOL64 = 1e-10


def _ctx(**kw):
    from qex_b200.engine import XCContext

    return XCContext(**kw)


def _to_np(t):
    return t.detach().cpu().numpy()


@pytest.fixture
def i8(monkeypatch):
    monkeypatch.setenv("QEXXC_I8", "1")


def test_mode_policy(monkeypatch):
    monkeypatch.delenv("QEXXC_I8", raising=False)
    assert _ctx(nao=120, ngrids_max=256, ncomp=1).contraction_mode == "dmma"
    assert _ctx(nao=256, ngrids_max=256, ncomp=1).contraction_mode == "int8"
    assert _ctx(nao=300, ngrids_max=256, ncomp=1, nbatch=2).contraction_mode == "dmma"  # batched contexts stay on DMMA
    monkeypatch.setenv("QEXXC_I8", "0")
    assert _ctx(nao=300, ngrids_max=256, ncomp=1).contraction_mode == "dmma"
    monkeypatch.setenv("QEXXC_I8", "1")
    assert _ctx(nao=8, ngrids_max=256, ncomp=1).contraction_mode == "int8"


@pytest.mark.parametrize("N,G,C", [(4, 1240, 1), (40, 1000, 1), (130, 3001, 4), (200, 515, 1), (70, 129, 4), (300, 4500, 1),
                                   (264, 2100, 4)])
@pytest.mark.parametrize("hermi", [0, 1])
def test_i8_eval_rho_and_vjp_vs_oracle(i8, N, G, C, hermi):
    ao, dm, w = synth_problem(N, G, C, seed=N + G)
    if hermi:
        dm = 0.5 * (dm + dm.transpose(0, 2, 1))
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=C)
    assert ctx.contraction_mode == "int8"
    ctx.set_grid(None, w).set_ao(ao, C)
    xct = "GGA" if C == 4 else "LDA"
    a = ao[0] if C == 4 else ao[0, 0]
    ref = numint_ref.eval_rho(a, dm[0], xct, hermi=hermi).reshape(C, G)
    got = _to_np(ctx.eval_rho(dm, ncomp=C, hermi=hermi))[0]
    assert rel_err(got, ref) <= TOL64
    rb = np.random.default_rng(1).standard_normal((C, G))
    ref_d = numint_ref.eval_rho_vjp(a, rb if C == 4 else rb[0], xct, hermi=hermi)
    got_d = _to_np(ctx.eval_rho_vjp(rb[None], ncomp=C, hermi=hermi))[0]
    assert rel_err(got_d, ref_d) <= TOL64


def test_i8_wide_dynamic_range(i8):
    """AO rows spanning 12 and columns spanning 6 orders of magnitude, per-point cotangents spanning 10 orders inside
    every 128-row sub-block, an all-zero column and all-zero rows: the fixed-point exponents must follow the data."""
    N, G = 150, 9000
    rng = np.random.default_rng(7)
    ao, dm, w = synth_problem(N, G, 1, seed=3)
    ao = ao * 10.0 ** rng.uniform(-12, 0, (1, 1, G, 1)) * 10.0 ** rng.uniform(-3, 3, (1, 1, 1, N))
    ao[..., 17] = 0.0
    ao[:, :, 4000:4200, :] = 0.0
    dm = 0.5 * (dm + dm.transpose(0, 2, 1))
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=1)
    ctx.set_grid(None, w).set_ao(ao, 1)
    ref = numint_ref.eval_rho(ao[0, 0], dm[0], "LDA", hermi=1).reshape(1, G)
    got = _to_np(ctx.eval_rho(dm, ncomp=1, hermi=1))[0]
    assert np.isfinite(got).all()
    assert rel_err(got, ref) <= TOL64
    # row-wise: every grid row carries its own exponent, so each rho[g] is accurate relative to its own row's scale
    scale = np.einsum("gi,ij,gj->g", np.abs(ao[0, 0]), np.abs(dm[0]), np.abs(ao[0, 0]))
    live = scale > 0
    assert (np.abs(got[0] - ref[0])[live] / scale[live]).max() <= 1e-11
    rb = (rng.standard_normal((1, G)) * 10.0 ** rng.uniform(-8, 2, (1, G)))
    ref_d = numint_ref.eval_rho_vjp(ao[0, 0], rb[0], "LDA", hermi=1)
    got_d = _to_np(ctx.eval_rho_vjp(rb[None], ncomp=1, hermi=1))[0]
    assert np.isfinite(got_d).all()
    assert rel_err(got_d, ref_d) <= TOL64


@pytest.mark.parametrize("cfg,G", [("c5", 20000), ("c5gga", 12000)])
def test_i8_nr_rks_fwd_vjp_matches_dmma(monkeypatch, cfg, G):
    """The whole fwd + VJP step at the headline AO count (N = 1000) on both tensor pipes: V_xc, E_xc, nelec, dm_bar and
    theta_bar agree to 1e-10 of the largest element (both are ~1e-12 from the exact result)."""
    from qex_b200 import workloads

    wl = workloads.make(cfg, ngrids=G)
    N = wl.nao
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("QEXXC_I8", mode)
        ctx = _ctx(nao=N, ngrids_max=G, ncomp=wl.ncomp, net=workloads.net_spec(wl))
        assert ctx.contraction_mode == ("int8" if mode == "1" else "dmma")
        ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights)
        ctx.eval_ao(1 if wl.ncomp == 4 else 0)
        out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, wl.xctype)
        bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, wl.xctype)
        res[mode] = (_to_np(out)[0].copy(), _to_np(bar).copy())
        ctx.close()
    (o0, b0), (o1, b1) = res["0"], res["1"]
    nn = N * N
    assert rel_err(o1[:nn], o0[:nn]) <= TOL64          # V_xc
    assert abs(o1[nn] - o0[nn]) <= 1e-9                 # E_xc (Ha)
    assert abs(o1[nn + 1] - o0[nn + 1]) <= 1e-9 * max(1.0, abs(o0[nn + 1]))  # nelec
    assert rel_err(b1[:nn], b0[:nn]) <= TOL64          # dm_bar
    assert rel_err(b1[nn:], b0[nn:]) <= TOL64          # theta_bar


def test_i8_bit_reproducible(i8):
    N, G = 260, 6000
    ao, dm, w = synth_problem(N, G, 1, seed=11)
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=1)
    ctx.set_grid(None, w).set_ao(ao, 1)
    rb = np.random.default_rng(2).standard_normal((1, 1, G))
    a = _to_np(ctx.eval_rho(dm, ncomp=1, hermi=0)).copy(), _to_np(ctx.eval_rho_vjp(rb, ncomp=1, hermi=0)).copy()
    b = _to_np(ctx.eval_rho(dm, ncomp=1, hermi=0)).copy(), _to_np(ctx.eval_rho_vjp(rb, ncomp=1, hermi=0)).copy()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_i8_regrid_reuses_context(i8):
    """A second, smaller grid on the same context: the geometry planes are rebuilt (stale planes would give the old rho)."""
    N = 96
    ctx = _ctx(nao=N, ngrids_max=3000, ncomp=1)
    for G, seed in ((3000, 1), (1111, 2)):
        ao, dm, w = synth_problem(N, G, 1, seed=seed)
        ctx.set_grid(None, w).set_ao(ao, 1)
        ref = numint_ref.eval_rho(ao[0, 0], dm[0], "LDA", hermi=0).reshape(1, G)
        got = _to_np(ctx.eval_rho(dm, ncomp=1, hermi=0))[0][:, :G]
        assert rel_err(got, ref) <= TOL64


@pytest.mark.parametrize("N,G,nmo", [(40, 1500, 7), (300, 3000, 150), (130, 2100, 65)])
def test_i8_mo_form_of_rho(i8, N, G, nmo):
    """rho from occupied orbitals (pyscf eval_rho2, the dms.mo_coeff branch at numint_legacy.py:527-545) on the INT8 pipe:
    rho = sum_k occ_k (ao C_k)^2 against the dm form of the oracle, including negative and zero occupations."""
    ao, _, w = synth_problem(N, G, 1, seed=N + nmo)
    rng = np.random.default_rng(nmo)
    Cm = rng.standard_normal((N, nmo)) / np.sqrt(N)
    occ = rng.uniform(0.2, 2.0, nmo)
    occ[::5] *= -1.0
    occ[3] = 0.0
    dm = (Cm * occ) @ Cm.T
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=1)
    assert ctx.contraction_mode == "int8"
    ctx.set_grid(None, w).set_ao(ao, 1)
    ref = numint_ref.eval_rho(ao[0, 0], dm, "LDA", hermi=1).reshape(G)
    got = _to_np(ctx.eval_rho_mo(Cm, occ))[0, 0]
    assert rel_err(got, ref) <= TOL64


@pytest.mark.parametrize("C", [1, 4])
def test_i8_nset_density_matrices_over_a_shared_ao_tensor(i8, C):
    """nset density matrices of one molecule (numint_legacy.py:141-156) on the INT8 pipe: one set of digit planes of the
    shared AO tensor, the contractions looped over the sets; every set equals its single-dm oracle result."""
    N, G, B = 140, 2300, 3
    ao, dm, w = synth_problem(N, G, C, seed=21)
    rng = np.random.default_rng(22)
    dms = np.stack([dm[0], 0.5 * dm[0] + 0.01 * rng.standard_normal((N, N)), -1.5 * dm[0].T])
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=C, nbatch=B, shared_ao=True)
    assert ctx.contraction_mode == "int8"
    ctx.set_grid(None, w).set_ao(ao, C)
    xct = "GGA" if C == 4 else "LDA"
    a = ao[0] if C == 4 else ao[0, 0]
    got = _to_np(ctx.eval_rho(dms, ncomp=C, hermi=0))
    rb = rng.standard_normal((B, C, G))
    got_d = _to_np(ctx.eval_rho_vjp(rb, ncomp=C, hermi=0))
    for b in range(B):
        ref = numint_ref.eval_rho(a, dms[b], xct, hermi=0).reshape(C, G)
        assert rel_err(got[b], ref) <= TOL64
        ref_d = numint_ref.eval_rho_vjp(a, rb[b] if C == 4 else rb[b, 0], xct, hermi=0)
        assert rel_err(got_d[b], ref_d) <= TOL64
