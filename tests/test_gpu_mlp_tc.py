"""GPU parity of the FP32 network path on the 5th-generation tensor cores (csrc/xc_mlp_tc.cu: tcgen05.mma kind::tf32
with the 3xTF32 split, accumulators in tensor memory) against the float64 oracle, at the north-star tolerance for
FP32 network outputs (1e-5 relative).  Every case goes through the C ABI (`precision="f32"` contexts route to the
tensor-core kernels for width <= 64 and <= 3 hidden layers)."""
import numpy as np
import pytest

from oracle import mlp_ref
from tests._util import rel_err

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


def _ctx(**kw):
    from qex_b200.engine import XCContext

    return XCContext(**kw)


def _net(F=1, L=3, H=64, act="tanh", out_transform=0):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    return NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=F, n_hidden=L, width=H, activation=act, precision="f32",
                   out_transform=out_transform)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("L,H,act,G", [(3, 64, "tanh", 2000), (1, 64, "tanh", 333), (2, 32, "tanh", 129),
                                       (3, 17, "tanh", 1000), (3, 64, "gelu", 700), (2, 48, "softplus", 515),
                                       (3, 64, "swish", 640)])
def test_tc_local_mlp_fwd_vjp(L, H, act, G):
    spec = mlp_ref.MLPSpec([1] + [H] * L + [1], act)
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    rng = np.random.default_rng(5)
    rho = np.abs(rng.standard_normal(G)) * 1.5
    rho[::50] = 0.0
    ctx = _ctx(nao=4, ngrids_max=G, net=_net(L=L, H=H, act=act))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, _ = ctx.xc_fwd(rho, theta, "NN")
    e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho)
    assert rel_err(_np(exc)[0], e_ref) <= TOL32
    assert rel_err(_np(vrho)[0], v_ref) <= TOL32
    eb, vb = rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho, eb, vb)
    assert rel_err(_np(rbar)[0, 0], r_ref) <= TOL32
    assert rel_err(_np(tbar), t_ref) <= 5 * TOL32  # sums over G points of float32 terms
    # bitwise reproducible (fixed-order reductions, accumulators in tensor memory)
    rbar2, tbar2 = ctx.xc_vjp(rho, theta, eb, vb, xctype="NN")
    assert np.array_equal(_np(tbar), _np(tbar2)) and np.array_equal(_np(rbar), _np(rbar2))


def test_tc_local_mlp_gga_features():
    G = 900
    spec = mlp_ref.MLPSpec([2, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 7))
    rng = np.random.default_rng(9)
    rho = rng.standard_normal((4, G))
    rho[0] = np.abs(rho[0])
    sigma = (rho[1:] ** 2).sum(0)
    feats = np.stack([rho[0], sigma])
    ctx = _ctx(nao=4, ngrids_max=G, ncomp=4, net=_net(F=2))
    ctx.set_grid(None, np.ones(G))
    exc, vrho, vgamma = ctx.xc_fwd(rho, theta, "GGA")
    e_ref, g_ref = mlp_ref.exc_and_grad_features(spec, theta, feats)
    assert rel_err(_np(exc)[0], e_ref) <= TOL32
    assert rel_err(_np(vrho)[0], g_ref[0]) <= TOL32
    assert rel_err(_np(vgamma)[0], g_ref[1]) <= TOL32
    eb, vb, gb = rng.standard_normal(G), rng.standard_normal(G), rng.standard_normal(G)
    rbar, tbar = ctx.xc_vjp(rho, theta, eb, vb, gb, xctype="GGA")
    fb, t_ref = mlp_ref.exc_and_grad_features_vjp(spec, theta, feats, eb, np.stack([vb, gb]))
    r_ref = np.zeros((4, G))
    r_ref[0] = fb[0]
    r_ref[1:] = fb[1] * 2.0 * rho[1:]
    assert rel_err(_np(rbar)[0], r_ref) <= TOL32
    assert rel_err(_np(tbar), t_ref) <= 5 * TOL32


def test_tc_out_transform_and_batches():
    """The trainer's flax MLP wraps the output as -scale * swish(u) (trainer_legacy_no_jit.py:96-107); batched
    contexts (config c4) run every molecule's points through the same persistent CTAs."""
    B, G = 3, 450
    spec = mlp_ref.MLPSpec([1, 64, 64, 1], "gelu", out_transform="neg_scale_swish")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 2))
    rng = np.random.default_rng(1)
    rho = np.abs(rng.standard_normal((B, G)))
    ctx = _ctx(nao=4, ngrids_max=G, nbatch=B, net=_net(L=2, act="gelu", out_transform=1))
    ctx.set_grid(None, np.ones((B, G)))
    exc, vrho, _ = ctx.xc_fwd(rho[:, None, :], theta, "NN")
    eb, vb = rng.standard_normal((B, G)), rng.standard_normal((B, G))
    rbar, tbar = ctx.xc_vjp(rho[:, None, :], theta, eb, vb, xctype="NN")
    t_sum = 0
    for b in range(B):
        e_ref, v_ref = mlp_ref.exc_and_vrho_local(spec, theta, rho[b])
        assert rel_err(_np(exc)[b], e_ref) <= TOL32
        assert rel_err(_np(vrho)[b], v_ref) <= TOL32
        r_ref, t_ref = mlp_ref.exc_and_vrho_local_vjp(spec, theta, rho[b], eb[b], vb[b])
        assert rel_err(_np(rbar)[b, 0], r_ref) <= TOL32
        t_sum = t_sum + t_ref
    assert rel_err(_np(tbar), t_sum) <= 5 * TOL32
