"""GPU parity: every stage of the hot path, through the C ABI, against the float64 CPU oracle.

Tolerances (BASELINE.json north_star): float64 path |dE_xc| <= 1e-9 Ha, V_xc / gradient
elements within 1e-10 relative (to the largest element of the array); float32 network / QNN
outputs within 1e-5 relative.
"""
import numpy as np
import pytest

from oracle import gto_ref, mlp_ref, numint_ref, qnn_ref
from tests._util import rel_err, synth_problem

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5


def _ctx(**kw):
    from qex_b200.engine import XCContext

    return XCContext(**kw)


def _mlp_net(F=1, L=3, H=64, act="tanh", prec="f64", kind=None, out_transform=0):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    return NetSpec(kind=kind or _lib.NET_LOCAL_MLP, n_features=F, n_hidden=L, width=H, activation=act,
                   precision=prec, out_transform=out_transform)


def _to_np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("deriv", [0, 1])
def test_eval_ao_matches_oracle(deriv):
    from qex_b200 import gto

    basis = [[0, (0.8, 1.0), (0.3, 0.5)], [0, (2.1, 1.0)], [1, (0.9, 1.0), (0.25, 0.4)], [2, (0.7, 1.0)], [3, (0.6, 1.0)]]
    mol = gto.Mole([(6, (0.1, -0.2, 0.3)), (8, (1.4, 0.6, -0.5)), (1, (-1.0, 0.9, 0.2))], basis=basis, unit="Bohr")
    rng = np.random.default_rng(0)
    G = 777
    coords = rng.uniform(-3, 3, (G, 3))
    ref = gto_ref.eval_ao(mol._atm, mol._bas, mol._env, coords, deriv)
    ncomp = 4 if deriv else 1
    ctx = _ctx(nao=mol.nao_nr(), ngrids_max=G, ncomp=ncomp)
    ctx.set_grid(coords, np.ones(G)).set_basis(mol._atm, mol._bas, mol._env).eval_ao(deriv)
    got = _to_np(ctx.get_ao(ncomp))[0]
    ref = ref.reshape(ncomp, G, -1)
    assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,G,C", [(4, 1240, 1), (40, 1000, 1), (120, 3001, 4), (200, 515, 1), (70, 129, 4)])
@pytest.mark.parametrize("hermi", [0, 1])
def test_eval_rho_and_vjp(N, G, C, hermi):
    ao, dm, w = synth_problem(N, G, C, seed=N + G)
    if hermi:
        dm = 0.5 * (dm + dm.transpose(0, 2, 1))
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=C)
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


# ------------------------------------------------------------------------------------------------
# stage 3: local MLP
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["tanh", "gelu", "softplus", "sigmoid", "elu", "selu", "swish", "relu", "leaky_relu"])
def test_local_mlp_fwd_vjp_f64(act):
    G = 1000
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], act)
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    rng = np.random.default_rng(5)
    rho = np.abs(rng.standard_normal(G)) * 1.5
    rho[::50] = 0.0
    ctx = _ctx(nao=4, ngrids_max=G, net=_mlp_net(act=act))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_to_np(exc)[0], e_ref) <= TOL64
    assert rel_err(_to_np(vrho)[0], v_ref) <= TOL64
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_to_np(rbar)[0, 0], r_ref) <= TOL64
    assert rel_err(_to_np(tbar), t_ref) <= TOL64


@pytest.mark.parametrize("L,H", [(1, 64), (2, 32), (3, 17), (3, 64)])
def test_local_mlp_shapes(L, H):
    G = 333
    spec = mlp_ref.MLPSpec([1] + [H] * L + [1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 1))
    rng = np.random.default_rng(2)
    rho = np.abs(rng.standard_normal(G))
    ctx = _ctx(nao=4, ngrids_max=G, net=_mlp_net(L=L, H=H))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_to_np(exc)[0], e_ref) <= TOL64
    assert rel_err(_to_np(vrho)[0], v_ref) <= TOL64
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_to_np(rbar)[0, 0], r_ref) <= TOL64
    assert rel_err(_to_np(tbar), t_ref) <= TOL64


def test_local_mlp_f32_within_1e5():
    G = 2000
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    rng = np.random.default_rng(5)
    rho = np.abs(rng.standard_normal(G)) * 1.5
    ctx = _ctx(nao=4, ngrids_max=G, net=_mlp_net(prec="f32"))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_to_np(exc)[0], e_ref) <= TOL32
    assert rel_err(_to_np(vrho)[0], v_ref) <= TOL32
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_to_np(rbar)[0, 0], r_ref) <= TOL32
    assert rel_err(_to_np(tbar), t_ref) <= 5 * TOL32  # sums over 2000 points of float32 terms


def test_local_mlp_gga_features():
    G = 900
    spec = mlp_ref.MLPSpec([2, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 7))
    rng = np.random.default_rng(9)
    rho = rng.standard_normal((4, G))
    rho[0] = np.abs(rho[0])
    sigma = (rho[1:] ** 2).sum(0)
    feats = np.stack([rho[0], sigma])
    ctx = _ctx(nao=4, ngrids_max=G, ncomp=4, net=_mlp_net(F=2))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, vgamma = ctx.xc_fwd(rho, theta, "GGA")
    e_ref, g_ref = mlp_ref.exc_and_grad_features(spec, theta, feats)
    assert rel_err(_to_np(exc)[0], e_ref) <= TOL64
    assert rel_err(_to_np(vrho)[0], g_ref[0]) <= TOL64
    assert rel_err(_to_np(vgamma)[0], g_ref[1]) <= TOL64
    eb, vb, gb = rng.standard_normal(G), rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, gb, xctype="GGA")
    fb, t_ref = mlp_ref.exc_and_grad_features_vjp(spec, theta, feats, eb, np.stack([vb, gb]))
    r_ref = np.zeros((4, G))
    r_ref[0] = fb[0]
    r_ref[1:] = fb[1] * 2.0 * rho[1:]
    assert rel_err(_to_np(rbar)[0], r_ref) <= TOL64
    assert rel_err(_to_np(tbar), t_ref) <= TOL64


def test_apply_fn_local_mlp_and_vjp():
    G = 513  # the 1D trainer's grid size
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 11))
    rng = np.random.default_rng(4)
    x = np.abs(rng.standard_normal(G))
    ctx = _ctx(nao=4, ngrids_max=G, net=_mlp_net())
    y = _to_np(ctx.apply_fn(x, theta))
    assert rel_err(y, mlp_ref.apply_local(spec, theta, x)) <= TOL64
    yb = rng.standard_normal(G)
    xb, tb = ctx.apply_fn_vjp(x, theta, yb)
    xb_ref, tb_ref = mlp_ref.apply_vjp(spec, theta, x, yb)
    assert rel_err(_to_np(xb), xb_ref[:, 0]) <= TOL64
    assert rel_err(_to_np(tb), tb_ref) <= TOL64


# ------------------------------------------------------------------------------------------------
# stage 3: global MLP and QNN
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,H,L,ot", [(1240, 64, 3, 0), (1192, 64, 3, 0), (300, 128, 2, 1)])
def test_global_mlp_fwd_vjp(G, H, L, ot):
    from qex_b200 import _lib

    spec = mlp_ref.MLPSpec([G] + [H] * L + [1], "gelu" if ot else "tanh", out_transform="neg_scale_swish" if ot else "none")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 2))
    rng = np.random.default_rng(6)
    B = 2
    rho = np.abs(rng.standard_normal((B, G)))
    net = _mlp_net(L=L, H=H, act=spec.activation, kind=_lib.NET_GLOBAL_MLP, out_transform=ot)
    ctx = _ctx(nao=4, ngrids_max=G, nbatch=B, net=net)
    ctx.set_grid(None, np.ones((B, G)))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN-AmplitudeEncoding")
    eb, vb = rng.standard_normal(B), rng.standard_normal((B, G))
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN-AmplitudeEncoding")
    t_sum = 0
    for b in range(B):
        e_ref, v_ref = mlp_ref.exc_and_vrho_global(spec, theta, rho[b])
        assert abs(_to_np(exc)[b] - e_ref) <= TOL64 * max(1.0, abs(e_ref))
        assert rel_err(_to_np(vrho)[b], v_ref) <= TOL64
        r_ref, t_ref = mlp_ref.exc_and_vrho_global_vjp(spec, theta, rho[b], eb[b], vb[b])
        assert rel_err(_to_np(rbar)[b, 0], r_ref) <= TOL64
        t_sum = t_sum + t_ref
    assert rel_err(_to_np(tbar), t_sum) <= TOL64


@pytest.mark.parametrize("nq,nl", [(6, 2), (2, 2), (4, 1), (5, 3), (7, 1)])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_qnn_fwd_vjp(nq, nl, prec):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    G = 300 if nq < 7 else 64
    spec = qnn_ref.QNNSpec(nq, nl)
    theta = qnn_ref.init_params(spec, 0) * 3.0
    rng = np.random.default_rng(8)
    rho = np.abs(rng.standard_normal(G)) * 1.2
    rho[::17] = 0.0
    net = NetSpec(kind=_lib.NET_LOCAL_QNN, n_features=1, n_hidden=nl, width=nq, precision=prec, in_scale=1.0)
    ctx = _ctx(nao=4, ngrids_max=G, net=net)
    ctx.set_grid(None, np.ones(G))
    tol = TOL64 if prec == "f64" else TOL32
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = qnn_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_to_np(exc)[0], e_ref) <= tol
    assert rel_err(_to_np(vrho)[0], v_ref) <= tol
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = qnn_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_to_np(rbar)[0, 0], r_ref) <= tol
    assert rel_err(_to_np(tbar), t_ref) <= (tol if prec == "f64" else 20 * tol)


# ------------------------------------------------------------------------------------------------
# stage 4 and the fused hot path
# ------------------------------------------------------------------------------------------------
def _toy_eval_xc(xc_code, rho, *args, **kwargs):
    """The reference's own test functional (tests/test_numint.py:96-103)."""
    exc = 0.01 * rho**2
    vrho = 0.01 * 2 * rho
    return exc, (vrho, None, None, None), None, None


@pytest.mark.parametrize("N,G", [(4, 1240), (2, 306), (120, 2000), (33, 129)])
def test_vxc_assemble_toy_functional(N, G):
    ao, dm, w = synth_problem(N, G, 1, seed=3)
    ctx = _ctx(nao=N, ngrids_max=G)
    ctx.set_grid(None, w).set_ao(ao, 1)
    rho = ctx.eval_rho(dm, 1, 0)
    r = _to_np(rho)[0, 0]
    exc, (vrho, *_), _, _ = _toy_eval_xc("NN", r)
    out = _to_np(ctx.vxc_assemble(rho, exc, vrho, xctype="NN"))[0]
    nelec, excsum, vmat = numint_ref.nr_rks(ao[0, 0], w[0], dm[0], _toy_eval_xc, "NN")
    assert rel_err(out[: N * N].reshape(N, N), vmat) <= TOL64
    assert abs(out[N * N] - excsum) <= 1e-9
    assert abs(out[N * N + 1] - nelec) <= 1e-9 * max(1.0, abs(nelec))
    # closed form of the same thing (einsum statement, scf_functions_masked.py:143-159)
    dms = 0.5 * (dm[0] + dm[0].T)
    v2, e2 = numint_ref.get_veff_xc_einsum(ao[0, 0], w[0], dms, _toy_eval_xc)
    assert rel_err(out[: N * N].reshape(N, N), v2) <= 1e-9
    assert abs(out[N * N] - e2) <= 1e-9


def _nr_rks_oracle(kind, spec, theta, ao, w, dm, e_bar, v_bar, hermi=0):
    if kind == "NN":
        def eval_xc(code, rho, **kw):
            e, v = mlp_ref.exc_and_vrho_local(spec, theta, rho)
            return e, (v, None, None, None), None, None

        fwd = numint_ref.nr_rks(ao, w, dm, eval_xc, "NN", hermi=hermi)
        D, tb = numint_ref.nr_rks_vjp(
            ao, w, dm, lambda r, p: mlp_ref.exc_and_vrho_local(spec, theta, r),
            lambda r, p, eb, vb: mlp_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb), e_bar, v_bar, "NN", hermi=hermi)
    elif kind == "NN-AmplitudeEncoding":
        def eval_xc(code, rho, **kw):
            e, v = mlp_ref.exc_and_vrho_global(spec, theta, rho)
            return e, (v, None, None, None), None, None

        fwd = numint_ref.nr_rks(ao, w, dm, eval_xc, kind, hermi=hermi)
        D, tb = numint_ref.nr_rks_vjp(
            ao, w, dm, lambda r, p: mlp_ref.exc_and_vrho_global(spec, theta, r),
            lambda r, p, eb, vb: mlp_ref.exc_and_vrho_global_vjp(spec, theta, r, eb, vb), e_bar, v_bar, kind, hermi=hermi)
    else:
        def eval_xc(code, rho, **kw):
            feats = np.stack([rho[0], (rho[1:4] ** 2).sum(0)])
            e, g = mlp_ref.exc_and_grad_features(spec, theta, feats)
            return e, (g[0], g[1], None, None), None, None

        def xc_fwd(feats, p):
            e, g = mlp_ref.exc_and_grad_features(spec, theta, feats)
            return e, g[0], g[1]

        def xc_vjp(feats, p, eb, vb, gb):
            fb, tb = mlp_ref.exc_and_grad_features_vjp(spec, theta, feats, eb, np.stack([vb, gb]))
            return (fb[0], fb[1]), tb

        fwd = numint_ref.nr_rks(ao, w, dm, eval_xc, "GGA", hermi=hermi)
        D, tb = numint_ref.nr_rks_vjp(ao, w, dm, xc_fwd, xc_vjp, e_bar, v_bar, "GGA", hermi=hermi)
    return fwd, D, tb


@pytest.mark.parametrize("kind,N,G,B", [
    ("NN", 4, 1240, 1), ("NN", 120, 2000, 1), ("NN", 40, 700, 3), ("NN", 200, 515, 1), ("NN", 257, 1100, 1),
    ("NN-AmplitudeEncoding", 4, 1240, 3), ("NN-AmplitudeEncoding", 4, 1192, 1),
    ("GGA", 120, 2000, 1), ("GGA", 10, 400, 2),
    # edges: one AO / one point, tile boundaries (BN = 32/64/128 +- 1), the c5 AO count at a small grid, ragged GGA
    ("NN", 1, 1, 1), ("NN", 33, 127, 2), ("NN", 64, 128, 1), ("NN", 129, 257, 1), ("NN", 1000, 300, 1),
    ("GGA", 65, 130, 1), ("GGA", 136, 300, 2), ("NN-AmplitudeEncoding", 7, 33, 2),
    # beyond the c5 AO count (11 column tiles, ragged last one) and a multi-tile GGA case
    ("NN", 1290, 260, 1), ("GGA", 520, 200, 1),
])
def test_nr_rks_fwd_and_vjp(kind, N, G, B):
    from qex_b200 import _lib

    C = 4 if kind == "GGA" else 1
    ao, dm, w = synth_problem(N, G, C, B=B, seed=11)
    rng = np.random.default_rng(12)
    if kind == "NN":
        spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
        net = _mlp_net()
    elif kind == "GGA":
        spec = mlp_ref.MLPSpec([2, 64, 64, 64, 1], "tanh")
        net = _mlp_net(F=2)
    else:
        spec = mlp_ref.MLPSpec([G, 64, 64, 64, 1], "tanh")
        net = _mlp_net(kind=_lib.NET_GLOBAL_MLP)
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 5))
    ctx = _ctx(nao=N, ngrids_max=G, ncomp=C, nbatch=B, net=net)
    ctx.set_grid(None, w).set_ao(ao, C)
    out, resid = ctx.nr_rks_fwd(dm, theta, kind, hermi=0)
    e_bar = rng.standard_normal(B)
    v_bar = rng.standard_normal((B, N, N))
    bar = _to_np(ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, kind, hermi=0))
    out = _to_np(out)
    t_sum = 0
    for b in range(B):
        a = ao[b] if C == 4 else ao[b, 0]
        (nelec, excsum, vmat), D, tb = _nr_rks_oracle(kind, spec, theta, a, w[b], dm[b], e_bar[b], v_bar[b])
        assert rel_err(out[b, : N * N].reshape(N, N), vmat) <= TOL64
        assert abs(out[b, N * N] - excsum) <= 1e-9
        assert abs(out[b, N * N + 1] - nelec) <= 1e-9 * max(1.0, abs(nelec))
        assert rel_err(bar[b * N * N : (b + 1) * N * N].reshape(N, N), D) <= TOL64
        t_sum = t_sum + tb
    assert rel_err(bar[B * N * N :], t_sum) <= TOL64


def test_nr_rks_qnn_h2_size():
    """Config c2: LocalQNN (6 qubits, 2 layers) on an H2-size problem."""
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    N, G = 4, 1240
    ao, dm, w = synth_problem(N, G, 1, seed=21)
    spec = qnn_ref.QNNSpec(6, 2)
    theta = qnn_ref.init_params(spec, 0)
    net = NetSpec(kind=_lib.NET_LOCAL_QNN, n_hidden=2, width=6, in_scale=1.0)
    ctx = _ctx(nao=N, ngrids_max=G, net=net)
    ctx.set_grid(None, w).set_ao(ao, 1)
    out, resid = ctx.nr_rks_fwd(dm, theta, "NN")
    rng = np.random.default_rng(3)
    e_bar, v_bar = rng.standard_normal(1), rng.standard_normal((1, N, N))
    bar = _to_np(ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, "NN"))

    def eval_xc(code, rho, **kw):
        e, v = qnn_ref.exc_and_vrho_local(spec, theta, rho)
        return e, (v, None, None, None), None, None

    nelec, excsum, vmat = numint_ref.nr_rks(ao[0, 0], w[0], dm[0], eval_xc, "NN")
    D, tb = numint_ref.nr_rks_vjp(
        ao[0, 0], w[0], dm[0], lambda r, p: qnn_ref.exc_and_vrho_local(spec, theta, r),
        lambda r, p, eb, vb: qnn_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb), e_bar[0], v_bar[0], "NN")
    out = _to_np(out)[0]
    assert rel_err(out[: N * N].reshape(N, N), vmat) <= TOL64
    assert abs(out[N * N] - excsum) <= 1e-9
    assert rel_err(bar[: N * N].reshape(N, N), D) <= TOL64
    assert rel_err(bar[N * N :], tb) <= TOL64


def test_end_to_end_from_basis_tables():
    """Stages 1-4 from libcint tables + coords (no AO upload): the drop-in data flow."""
    from qex_b200 import gen_grid, gto

    mol = gto.h2(0.74, "6-31g")
    grids = gen_grid.Grids(mol, n_rad=20, n_theta=8, n_phi=8).build()
    G, N = grids.size, mol.nao_nr()
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 5))
    rng = np.random.default_rng(1)
    Cm = rng.standard_normal((N, 1))
    dm = 2.0 * Cm @ Cm.T
    ctx = _ctx(nao=N, ngrids_max=G, net=_mlp_net())
    ctx.set_grid(grids.coords, grids.weights).set_basis(mol._atm, mol._bas, mol._env).eval_ao(0)
    out, _ = ctx.nr_rks_fwd(dm, theta, "NN")
    ao = gto_ref.eval_ao(mol._atm, mol._bas, mol._env, grids.coords, 0)

    def eval_xc(code, rho, **kw):
        e, v = mlp_ref.exc_and_vrho_local(spec, theta, rho)
        return e, (v, None, None, None), None, None

    nelec, excsum, vmat = numint_ref.nr_rks(ao, grids.weights, dm, eval_xc, "NN")
    out = _to_np(out)[0]
    assert rel_err(out[: N * N].reshape(N, N), vmat) <= TOL64
    assert abs(out[N * N] - excsum) <= 1e-9
    assert abs(out[N * N + 1] - nelec) <= 1e-9 * max(1.0, abs(nelec))


@pytest.mark.parametrize("N,G,nmo", [(4, 1240, 1), (120, 2000, 5), (200, 515, 37), (257, 1100, 150)])
def test_mo_form_of_rho_and_nr_rks(N, G, nmo):
    """NumInt._gen_rho_evaluator's mo_coeff branch (numint_legacy.py:527-545 -> eval_rho2): same rho and
    the same nr_rks outputs as the dense-dm path with dm = C occ C^T (one negative occupation included)."""
    rng = np.random.default_rng(N + nmo)
    ao = rng.standard_normal((G, N)) * np.exp(-np.abs(rng.standard_normal((G, 1))) * 1.5) * 0.6
    w = np.abs(rng.standard_normal(G)) * 10.0 / G
    Cm = rng.standard_normal((N, nmo)) / np.sqrt(N)
    occ = np.full(nmo, 2.0)
    occ[-1] = 1.0
    if nmo > 2:
        occ[1] = -0.5
        occ[2] = 0.0  # dropped (|occ| <= OCCDROP)
    dm = (Cm * occ) @ Cm.T
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 2))
    ctx = _ctx(nao=N, ngrids_max=G, net=_mlp_net())
    ctx.set_grid(None, w).set_ao(ao, 1)
    rho = _to_np(ctx.eval_rho_mo(Cm, occ))[0, 0]
    assert rel_err(rho, numint_ref.eval_rho2(ao, Cm, occ)) <= TOL64
    assert rel_err(rho, numint_ref.eval_rho(ao, dm, "LDA")) <= 1e-9  # the two forms agree up to rounding
    out_mo, resid = ctx.nr_rks_fwd_mo(Cm, occ, theta, "NN")
    out_dm, _ = ctx.nr_rks_fwd(dm, theta, "NN")
    assert rel_err(_to_np(out_mo), _to_np(out_dm)) <= 1e-9
    e_bar, v_bar = rng.standard_normal(1), rng.standard_normal((1, N, N))
    bar = _to_np(ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, "NN"))
    D, tb = numint_ref.nr_rks_vjp(
        ao, w, dm, lambda r, p: mlp_ref.exc_and_vrho_local(spec, theta, r),
        lambda r, p, eb, vb: mlp_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb), e_bar[0], v_bar[0], "NN")
    assert rel_err(bar[: N * N].reshape(N, N), D) <= 1e-9
    assert rel_err(bar[N * N :], tb) <= 1e-9


def test_errors_are_loud():
    from qex_b200 import _lib

    ctx = _ctx(nao=4, ngrids_max=128, net=_mlp_net())
    with pytest.raises(_lib.QexxcError):
        ctx.nr_rks_fwd(np.eye(4), np.zeros(ctx.n_params), "NN")  # no grid / AO yet
    # any width / depth is served (csrc/xc_mlp_wide.cu beyond 64 x 3); what the kernels do not have is more than the
    # two input features (rho, sigma) the reference's local functionals use
    with pytest.raises(NotImplementedError):
        _ctx(nao=4, ngrids_max=128, ncomp=4, net=_mlp_net(F=3)).set_grid(None, np.ones(128)).xc_fwd(
            np.ones((4, 128)), np.zeros(3 * 64 + 64 + 2 * (64 * 64 + 64) + 65), "GGA")
    with pytest.raises(NotImplementedError):
        _ctx(nao=4, ngrids_max=128, net=_mlp_net(F=3, H=100))


def test_empty_grid_and_zero_density_edge_cases():
    """An empty grid (the reference's block loop simply does not run: nelec = excsum = 0, vmat = 0) and an
    all-zero density matrix (rho = 0 exactly at every point) go through the same kernels without special cases."""
    N = 12
    ctx = _ctx(nao=N, ngrids_max=256, net=_mlp_net())
    theta = mlp_ref.pack(*mlp_ref.init_params(mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh"), 1))
    rng = np.random.default_rng(0)
    dm = rng.standard_normal((N, N))
    ctx.set_grid(None, np.zeros(0)).set_ao(np.zeros((1, 1, 0, N)), 1)
    out, resid = ctx.nr_rks_fwd(dm, theta, "NN")
    assert not _to_np(out).any()
    bar = _to_np(ctx.nr_rks_vjp(theta, resid, [1.0], rng.standard_normal((N, N)), "NN"))
    assert not bar[: N * N].any() and np.isfinite(bar).all()
    # zero density on a real grid: exc = f(0) per point, nelec = 0, and the oracle agrees
    ao, _, w = synth_problem(N, 200, 1, seed=3)
    ctx.set_grid(None, w).set_ao(ao, 1)
    out = _to_np(ctx.nr_rks_fwd(np.zeros((N, N)), theta, "NN")[0])[0]
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    nelec, excsum, vmat = numint_ref.nr_rks(ao[0, 0], w[0], np.zeros((N, N)),
                                            lambda code, rho, **k: (mlp_ref.exc_and_vrho_local(spec, theta, rho)[0],
                                                                    (mlp_ref.exc_and_vrho_local(spec, theta, rho)[1], None, None, None), None, None),
                                            "NN")
    assert out[N * N + 1] == 0.0 and abs(out[N * N] - excsum) <= 1e-12
    assert rel_err(out[: N * N].reshape(N, N), vmat) <= TOL64
