"""CPU tests that pin the oracle (no GPU): closed forms, the reference's own known-answer
conventions, an independent autograd (torch float64), finite differences.

The reference cannot be imported here (no jax/pyscf/horqrux), and its tests hold no golden
vectors for this path, so these are the anchors the parity claim rests on.
"""
import numpy as np
import pytest
import torch

from oracle import gto_ref, mlp_ref, numint_ref, qnn_ref
from qex_b200 import gen_grid, gto
from tests._util import rel_err, synth_problem


# ---- AO evaluation ---------------------------------------------------------------------------
def test_ao_orthonormal_spdf():
    """pyscf AOs are normalised, real solid harmonics orthogonal: numerical overlap = identity."""
    basis = [[0, (0.8, 1.0), (0.3, 0.5)], [1, (0.9, 1.0)], [2, (0.7, 1.0), (1.9, 0.3)], [3, (0.6, 1.0)]]
    m = gto.Mole([(6, (0.1, -0.2, 0.3))], basis=basis, unit="Bohr")
    g = gen_grid.Grids(m, n_rad=80, n_theta=24, n_phi=24).build()
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    S = ao.T @ (ao * g.weights[:, None])
    assert m.nao_nr() == 1 + 3 + 5 + 7
    assert np.abs(S - np.eye(16)).max() < 1e-10


def test_ao_gradient_finite_difference():
    basis = [[0, (0.8, 1.0), (0.3, 0.5)], [1, (0.9, 1.0)], [2, (0.7, 1.0)], [3, (0.6, 1.0)]]
    m = gto.Mole([(6, (0.1, -0.2, 0.3)), (6, (1.1, 0.7, -0.4))], basis=basis, unit="Bohr")
    pts = np.random.default_rng(0).uniform(-2, 2, (100, 3))
    ao1 = gto_ref.eval_ao(m._atm, m._bas, m._env, pts, 1)
    h = 1e-6
    for k in range(3):
        d = np.zeros(3)
        d[k] = h
        fd = (gto_ref.eval_ao(m._atm, m._bas, m._env, pts + d, 0) - gto_ref.eval_ao(m._atm, m._bas, m._env, pts - d, 0)) / (2 * h)
        assert np.abs(fd - ao1[k + 1]).max() < 1e-8


def test_h2_631g_tables_and_electron_count():
    """H2/6-31G (the README molecule): 4 AOs, 2 s-shells per atom; a normalised doubly occupied
    bonding orbital integrates to 2 electrons on the grid."""
    m = gto.h2(0.74, "6-31g")
    assert m.nao_nr() == 4 and m.nbas == 4
    assert np.allclose(m.atom_coords()[1, 2], 0.74 / gto.BOHR)
    g = gen_grid.Grids(m, n_rad=60, n_theta=16, n_phi=16).build()
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    S = ao.T @ (ao * g.weights[:, None])
    c = np.array([1.0, 0.5, 1.0, 0.5])
    c = c / np.sqrt(c @ S @ c)
    dm = 2.0 * np.outer(c, c)
    rho = numint_ref.eval_rho(ao, dm, "LDA")
    assert abs(np.dot(rho, g.weights) - 2.0) < 1e-9
    # reference known-answer (tests/test_make_rdm1.py:85-97): identity MOs, occ [2, 0] -> diag(2, 0)
    assert np.allclose((np.eye(2) * np.array([2.0, 0.0])) @ np.eye(2).T, np.diag([2.0, 0.0]))


# ---- contractions ----------------------------------------------------------------------------
def _toy(xc_code, rho, *a, **k):
    """tests/test_numint.py:96-103 of the reference."""
    return 0.01 * rho**2, (0.02 * rho, None, None, None), None, None


@pytest.mark.parametrize("N,G", [(4, 1240), (7, 300)])
def test_nr_rks_closed_form_toy_functional(N, G):
    ao, dm, w = synth_problem(N, G, 1, seed=1)
    ao, dm, w = ao[0, 0], dm[0], w[0]
    nelec, exc, vmat = numint_ref.nr_rks(ao, w, dm, _toy, "NN")
    dms = 0.5 * (dm + dm.T)
    rho = np.einsum("gi,ij,gj->g", ao, dms, ao)
    assert abs(nelec - np.dot(rho, w)) < 1e-12
    assert abs(exc - np.sum(w * rho * 0.01 * rho**2)) < 1e-12
    V = np.einsum("gi,g,gj->ij", ao, w * 0.02 * rho, ao)  # 0.5*w*vrho then + transpose == w*vrho
    assert rel_err(vmat, V) < 1e-13
    v2, e2 = numint_ref.get_veff_xc_einsum(ao, w, dms, _toy)
    assert rel_err(vmat, v2) < 1e-13 and abs(exc - e2) < 1e-12
    # global branch: excsum is the functional's scalar, no weights (numint_legacy.py:331)
    def glob(code, rho, **k):
        return np.sum(0.01 * rho**2), (0.02 * rho, None, None, None), None, None

    _, exc_g, vmat_g = numint_ref.nr_rks(ao, w, dm, glob, "NN-AmplitudeEncoding")
    assert abs(exc_g - np.sum(0.01 * rho**2)) < 1e-12
    assert rel_err(vmat_g, V) < 1e-13


def test_blocked_accumulation_is_block_size_independent():
    ao, dm, w = synth_problem(6, 1000, 1, seed=2)
    a = numint_ref.nr_rks(ao[0, 0], w[0], dm[0], _toy, "NN", blksize=128)
    b = numint_ref.nr_rks(ao[0, 0], w[0], dm[0], _toy, "NN", blksize=1000)
    assert abs(a[1] - b[1]) < 1e-13 and rel_err(a[2], b[2]) < 1e-13


def test_eval_rho_gga_convention():
    """rho[0] = <c0, ao0>, rho[k] = 2 <c0, ao_k>  (numint_legacy.py:401-410)."""
    ao, dm, _ = synth_problem(5, 64, 4, seed=3)
    dms = 0.5 * (dm[0] + dm[0].T)
    rho = numint_ref.eval_rho(ao[0], dm[0], "GGA")
    assert rel_err(rho[0], np.einsum("gi,ij,gj->g", ao[0, 0], dms, ao[0, 0])) < 1e-13
    for k in (1, 2, 3):
        assert rel_err(rho[k], 2 * np.einsum("gi,ij,gj->g", ao[0, k], dms, ao[0, 0])) < 1e-13


@pytest.mark.parametrize("kind", ["NN", "GGA"])
def test_nr_rks_vjp_finite_difference(kind):
    N, G = 5, 60
    C = 4 if kind == "GGA" else 1
    ao, dm, w = synth_problem(N, G, C, seed=4)
    a = ao[0] if C == 4 else ao[0, 0]
    F = 2 if C == 4 else 1
    spec = mlp_ref.MLPSpec([F, 8, 8, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
    rng = np.random.default_rng(0)
    e_bar, v_bar = 0.7, rng.standard_normal((N, N))

    def loss(dm_, th_):
        if C == 1:
            def exc_fn(code, rho, **k):
                e, v = mlp_ref.exc_and_vrho_local(spec, th_, rho)
                return e, (v, None, None, None), None, None
        else:
            def exc_fn(code, rho, **k):
                e, g = mlp_ref.exc_and_grad_features(spec, th_, np.stack([rho[0], (rho[1:4] ** 2).sum(0)]))
                return e, (g[0], g[1], None, None), None, None
        _, e, v = numint_ref.nr_rks(a, w[0], dm_, exc_fn, kind)
        return e_bar * e + np.sum(v_bar * v)

    if C == 1:
        fwd = lambda r, p: mlp_ref.exc_and_vrho_local(spec, theta, r)
        vjp = lambda r, p, eb, vb: mlp_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb)
    else:
        def fwd(f, p):
            e, g = mlp_ref.exc_and_grad_features(spec, theta, f)
            return e, g[0], g[1]

        def vjp(f, p, eb, vb, gb):
            fb, tb = mlp_ref.exc_and_grad_features_vjp(spec, theta, f, eb, np.stack([vb, gb]))
            return (fb[0], fb[1]), tb
    D, tb = numint_ref.nr_rks_vjp(a, w[0], dm[0], fwd, vjp, e_bar, v_bar, kind)
    h = 1e-6
    for _ in range(6):
        i, j = rng.integers(0, N, 2)
        d = np.zeros((N, N))
        d[i, j] = h
        fd = (loss(dm[0] + d, theta) - loss(dm[0] - d, theta)) / (2 * h)
        assert abs(fd - D[i, j]) < 1e-7 * max(1.0, abs(fd))
    for k in rng.integers(0, theta.size, 6):
        d = np.zeros_like(theta)
        d[k] = h
        fd = (loss(dm[0], theta + d) - loss(dm[0], theta - d)) / (2 * h)
        assert abs(fd - tb[k]) < 1e-7 * max(1.0, abs(fd))


# ---- MLP -------------------------------------------------------------------------------------
_TORCH_ACT = {
    "tanh": torch.tanh, "relu": torch.relu, "softplus": torch.nn.functional.softplus, "sigmoid": torch.sigmoid,
    "elu": torch.nn.functional.elu, "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.01),
    "selu": torch.selu, "gelu": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
    "swish": torch.nn.functional.silu,
}


def _torch_mlp(spec, theta, X):
    Ws, bs, o = [], [], 0
    for fi, fo in zip(spec.sizes[:-1], spec.sizes[1:]):
        Ws.append(theta[o : o + fi * fo].reshape(fi, fo))
        bs.append(theta[o + fi * fo : o + fi * fo + fo])
        o += fi * fo + fo
    h = X * spec.in_scale
    for l, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ W + b
        if l < len(Ws) - 1:
            h = _TORCH_ACT[spec.activation](h)
    if spec.out_transform == "neg_scale_swish":
        h = -spec.out_scale * torch.nn.functional.silu(h)
    return h.sum(1)


@pytest.mark.parametrize("act", list(mlp_ref.ACTIVATIONS))
@pytest.mark.parametrize("F", [1, 2])
def test_mlp_against_torch_double_autograd(act, F):
    """stax Dense/activation semantics and the hand-written second-order VJP vs torch float64."""
    spec = mlp_ref.MLPSpec([F, 16, 16, 1], act, out_transform="neg_scale_swish" if act == "gelu" else "none")
    th = mlp_ref.pack(*mlp_ref.init_params(spec, 1))
    rng = np.random.default_rng(2)
    X = rng.standard_normal((9, F)) + 0.3
    tt = torch.tensor(th, dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    y = _torch_mlp(spec, tt, Xt)
    (g,) = torch.autograd.grad(y.sum(), Xt, create_graph=True)
    Ws, bs = mlp_ref.unpack(spec, th)
    y_ref, g_ref = mlp_ref.value_and_grad_x(spec, Ws, bs, X)
    assert rel_err(y_ref, y.detach().numpy()) < 1e-13
    assert rel_err(g_ref, g.detach().numpy()) < 1e-12
    yb, gb = rng.standard_normal(9), rng.standard_normal((9, F))
    L = (torch.tensor(yb) * y).sum() + (torch.tensor(gb) * g).sum()
    xb_t, tb_t = torch.autograd.grad(L, [Xt, tt])
    xb, tb = mlp_ref.second_order_vjp(spec, th, X, yb, gb)
    assert rel_err(xb, xb_t.numpy()) < 1e-11
    assert rel_err(tb, tb_t.numpy()) < 1e-11


def test_mlp_local_and_global_entry_points():
    spec = mlp_ref.MLPSpec([1, 8, 8, 1], "tanh")
    th = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
    rho = np.abs(np.random.default_rng(0).standard_normal(11))
    # apply_fn accepts [G] and [G,1] and divides by density_normalization_factor = 2
    assert np.allclose(mlp_ref.apply_local(spec, th, rho), mlp_ref.apply_local(spec, th, rho[:, None]))
    Ws, bs = mlp_ref.unpack(spec, th)
    h = np.tanh((rho[:, None] / 2.0) @ Ws[0] + bs[0])
    h = np.tanh(h @ Ws[1] + bs[1])
    assert np.allclose(mlp_ref.apply_local(spec, th, rho), (h @ Ws[2] + bs[2])[:, 0])
    # stax parameter list round trip
    Ws2, bs2 = mlp_ref.from_stax(mlp_ref.to_stax(Ws, bs))
    assert all(np.array_equal(a, b) for a, b in zip(Ws + bs, Ws2 + bs2))
    gspec = mlp_ref.MLPSpec([11, 8, 1], "tanh")
    gth = mlp_ref.pack(*mlp_ref.init_params(gspec, 0))
    e, v = mlp_ref.exc_and_vrho_global(gspec, gth, rho)
    hh = 1e-6
    fd = np.array([(mlp_ref.apply_global(gspec, gth, rho + hh * np.eye(11)[i]) -
                    mlp_ref.apply_global(gspec, gth, rho - hh * np.eye(11)[i]))[0] / (2 * hh) for i in range(11)])
    assert np.abs(fd - v).max() < 1e-8
    assert np.isclose(e, mlp_ref.apply_global(gspec, gth, rho).sum())


# ---- QNN -------------------------------------------------------------------------------------
def test_qnn_known_answer_conventions():
    """tests/test_measurements.py:31-61 and tests/test_quantum_measurement.py:46-59 of the reference."""
    s = np.zeros(4, complex)
    s[0] = 1
    assert np.allclose(qnn_ref.per_qubit_z(s, 2), [1, 1])           # |00> -> [1, 1], total 2
    assert np.allclose(qnn_ref.per_qubit_z(qnn_ref.apply_1q(s, qnn_ref._X, 0, 2), 2), [-1, 1])  # X(0)
    t = np.array([[0, 1], [0, 1]], complex).reshape(-1) / np.sqrt(2)   # state [[0,1],[0,1]]
    z = qnn_ref.per_qubit_z(t, 2)
    assert np.allclose(z * 2, [0, -2])  # unnormalised state of the reference test gives [0, -2]
    # rotation matrices: R_P(t) = cos(t/2) I - i sin(t/2) P
    assert np.allclose(qnn_ref.rot("Y", np.pi) @ [1, 0], [0, 1])
    assert np.allclose(qnn_ref.rot("X", np.pi) @ [1, 0], [0, -1j])


def _heisenberg_z_supports(n):
    """Propagate Z_i backwards through hea's CNOT block (n rings of CNOT(i -> i+1)) over GF(2):
    under CNOT(c,t), Z_t -> Z_c Z_t and Z_c -> Z_c.  Returns for every output qubit the set of
    input qubits whose Z-parity it measures."""
    gates = [(i, (i + 1) % n) for _ in range(n) for i in range(n)]
    supports = []
    for q in range(n):
        s = {q}
        for c, t in reversed(gates):
            if t in s:
                s ^= {c}
        supports.append(s)
    return supports


@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_qnn_zero_theta_closed_form(n):
    """theta = 0: the state is prod RY(x)|0>, a product state with <Z_j> = cos x; the CNOT block is
    a basis permutation, so sum_i <Z_i> = sum_i cos(x)^{|S_i|} with S_i from GF(2) propagation."""
    spec = qnn_ref.QNNSpec(n, 1)
    x = np.array([0.0, 0.4, 1.3, 2.9])
    got = qnn_ref.apply(spec, np.zeros(spec.n_params()), x)
    sup = _heisenberg_z_supports(n)
    want = sum(np.cos(x) ** len(s) for s in sup)
    assert np.abs(got - want).max() < 1e-12


def test_qnn_derivatives_finite_difference():
    spec = qnn_ref.QNNSpec(4, 2)
    th = qnn_ref.init_params(spec, 0) * 5
    x = np.array([0.0, 0.3, 1.1])
    e, v = qnn_ref.exc_and_vrho_local(spec, th, x)
    h = 1e-5
    fd = (qnn_ref.apply(spec, th, x + h) - qnn_ref.apply(spec, th, x - h)) / (2 * h)
    assert np.abs(fd - v).max() < 1e-8
    eb, vb = np.array([0.3, -1.0, 0.5]), np.array([1.0, 0.2, -0.7])
    rb, tb = qnn_ref.exc_and_vrho_local_vjp(spec, th, x, eb, vb)

    def L(th_, x_):
        e_, v_ = qnn_ref.exc_and_vrho_local(spec, th_, x_)
        return eb @ e_ + vb @ v_

    for k in range(0, spec.n_params(), 5):
        d = np.zeros_like(th)
        d[k] = h
        assert abs((L(th + d, x) - L(th - d, x)) / (2 * h) - tb[k]) < 1e-7
    for i in range(3):
        d = np.zeros(3)
        d[i] = h
        assert abs((L(th, x + d) - L(th, x - d)) / (2 * h) - rb[i]) < 1e-7
    assert spec.n_params() == 3 * 4 * 2 and len(qnn_ref.ansatz_gates(spec)) == 3 * 4 * 2 + 2
