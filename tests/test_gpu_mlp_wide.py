"""GPU parity of LocalMLP beyond 64 x 3 (csrc/xc_mlp_wide.cu): the reference takes any n_neurons / n_layers
(qedft/models/networks.py:103-108) and its 3D trainer defaults to a 512-wide gelu MLP with the -scale*swish output
(trainer_legacy_no_jit.py:96-107,136-140).  Float64 tolerances, through the C ABI."""
import numpy as np
import pytest

from oracle import mlp_ref
from tests._util import rel_err

pytestmark = pytest.mark.gpu
TOL64 = 1e-10


def _ctx(**kw):
    from qex_b200.engine import XCContext

    return XCContext(**kw)


def _net(F=1, L=3, H=64, act="tanh", out_transform=0, prec="f64"):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    return NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=F, n_hidden=L, width=H, activation=act, precision=prec,
                   out_transform=out_transform)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("L,H,act,ot,G", [(2, 512, "gelu", 1, 700), (3, 128, "tanh", 0, 333), (5, 64, "tanh", 0, 1000),
                                          (4, 100, "softplus", 0, 515), (1, 200, "swish", 0, 129), (2, 96, "elu", 0, 64)])
def test_wide_local_mlp_fwd_vjp(L, H, act, ot, G):
    spec = mlp_ref.MLPSpec([1] + [H] * L + [1], act, out_transform="neg_scale_swish" if ot else "none")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    rng = np.random.default_rng(5)
    rho = np.abs(rng.standard_normal(G)) * 1.5
    rho[::50] = 0.0
    ctx = _ctx(nao=4, ngrids_max=G, net=_net(L=L, H=H, act=act, out_transform=ot))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_np(exc)[0], e_ref) <= TOL64
    assert rel_err(_np(vrho)[0], v_ref) <= TOL64
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_np(rbar)[0, 0], r_ref) <= TOL64
    assert rel_err(_np(tbar), t_ref) <= TOL64
    rbar2, tbar2 = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    assert np.array_equal(_np(tbar), _np(tbar2)) and np.array_equal(_np(rbar), _np(rbar2))


def test_wide_gga_features_and_chunks():
    """Two input features (rho, sigma) and a grid larger than one chunk of the layer-by-layer path."""
    G = 40_000
    spec = mlp_ref.MLPSpec([2, 128, 128, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 7))
    rng = np.random.default_rng(9)
    rho = rng.standard_normal((4, G))
    rho[0] = np.abs(rho[0])
    sigma = (rho[1:] ** 2).sum(0)
    feats = np.stack([rho[0], sigma])
    ctx = _ctx(nao=4, ngrids_max=G, ncomp=4, net=_net(F=2, L=2, H=128))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, vgamma = ctx.xc_fwd(rho, theta, "GGA")
    e_ref, g_ref = mlp_ref.exc_and_grad_features(spec, theta, feats)
    assert rel_err(_np(exc)[0], e_ref) <= TOL64
    assert rel_err(_np(vrho)[0], g_ref[0]) <= TOL64
    assert rel_err(_np(vgamma)[0], g_ref[1]) <= TOL64
    eb, vb, gb = rng.standard_normal(G), rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, gb, xctype="GGA")
    fb, t_ref = mlp_ref.exc_and_grad_features_vjp(spec, theta, feats, eb, np.stack([vb, gb]))
    r_ref = np.zeros((4, G))
    r_ref[0] = fb[0]
    r_ref[1:] = fb[1] * 2.0 * rho[1:]
    assert rel_err(_np(rbar)[0], r_ref) <= TOL64
    assert rel_err(_np(tbar), t_ref) <= TOL64


def test_trainer_default_512_wide_network_in_nr_rks():
    """`nr_rks` forward + VJP with the trainer's default network shape (1 -> 512 -> 512 -> 1, gelu, -scale*swish) on a
    small synthetic molecule, against the oracle step."""
    from oracle import step_ref
    from qex_b200 import workloads
    from qex_b200.engine import XCContext

    wl = workloads.make("c3", ngrids=1500)
    net = dict(kind="local_mlp", n_features=1, n_hidden=2, width=512, activation="gelu", out_transform=1)
    wl.net, wl.xctype, wl.ncomp = net, "NN", 1
    wl.theta = workloads._mlp_theta([1, 512, 512, 1], 4)
    N, G = wl.nao, wl.ngrids
    ctx = XCContext(nao=N, ngrids_max=G, ncomp=1, net=workloads.net_spec(wl))
    ctx.set_basis(wl.mol._atm, wl.mol._bas, wl.mol._env).set_grid(wl.coords, wl.weights).eval_ao(0)
    out, resid = ctx.nr_rks_fwd(wl.dm, wl.theta, "NN")
    bar = ctx.nr_rks_vjp(wl.theta, resid, [wl.e_bar], wl.v_bar, "NN")
    out, bar = _np(out)[0], _np(bar)
    m = wl.mol
    ref = step_ref.xc_step(m._atm, m._bas, m._env, wl.coords, wl.weights, wl.dm, wl.net, wl.theta, "NN", wl.e_bar, wl.v_bar)
    assert rel_err(out[: N * N].reshape(N, N), ref["vmat"]) <= TOL64
    assert abs(out[N * N] - ref["excsum"]) <= 1e-9
    assert rel_err(bar[: N * N].reshape(N, N), ref["dm_bar"]) <= TOL64
    assert rel_err(bar[N * N:], ref["theta_bar"]) <= TOL64
