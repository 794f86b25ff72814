"""Oracle: one full XC step (AO evaluation + nr_rks forward + its VJP) on the CPU.
TEST INFRASTRUCTURE ONLY -- also the timed body of bench.py's CPU baseline / ``--impl reference``.

Glue over gto_ref / numint_ref / mlp_ref / qnn_ref in the order the reference executes them:
eval_ao inside block_loop (numint_legacy.py:292), make_rho (:294), eval_xc (:295-303), stage 4
(:304-309, 336-337), then the reverse pass JAX would run (trainer_legacy_no_jit.py:284).
"""
from __future__ import annotations

import numpy as np

from . import gto_ref, mlp_ref, numint_ref, qnn_ref


def _functional(net: dict, theta, ngrids):
    kind = net["kind"]
    if kind == "local_mlp":
        F = net.get("n_features", 1)
        spec = mlp_ref.MLPSpec([F] + [net["width"]] * net["n_hidden"] + [1], net.get("activation", "tanh"),
                               in_scale=net.get("in_scale", 0.5),
                               out_transform="neg_scale_swish" if net.get("out_transform") else "none")
        if F == 1:
            fwd = lambda r, p: mlp_ref.exc_and_vrho_local(spec, theta, r)
            vjp = lambda r, p, eb, vb: mlp_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb)
        else:
            def fwd(feats, p):
                e, g = mlp_ref.exc_and_grad_features(spec, theta, feats)
                return e, g[0], g[1]

            def vjp(feats, p, eb, vb, gb):
                fb, tb = mlp_ref.exc_and_grad_features_vjp(spec, theta, feats, eb, np.stack([vb, gb]))
                return (fb[0], fb[1]), tb
        return fwd, vjp
    if kind == "local_qnn":
        spec = qnn_ref.QNNSpec(net["width"], net["n_hidden"])
        return (lambda r, p: qnn_ref.exc_and_vrho_local(spec, theta, r),
                lambda r, p, eb, vb: qnn_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb))
    if kind == "global_mlp":
        spec = mlp_ref.MLPSpec([ngrids] + [net["width"]] * net["n_hidden"] + [1], net.get("activation", "tanh"),
                               in_scale=net.get("in_scale", 0.5))
        return (lambda r, p: mlp_ref.exc_and_vrho_global(spec, theta, r),
                lambda r, p, eb, vb: mlp_ref.exc_and_vrho_global_vjp(spec, theta, r, eb, vb))
    raise ValueError(kind)


def xc_step(atm, bas, env, coords, weights, dm, net, theta, xctype, e_bar, v_bar, hermi=0):
    """-> dict(nelec, excsum, vmat, dm_bar, theta_bar).  One forward + one reverse pass."""
    gga = xctype == "GGA"
    ao = gto_ref.eval_ao(atm, bas, env, coords, 1 if gga else 0)
    fwd, vjp = _functional(net, theta, coords.shape[0])

    def eval_xc(code, rho, **kw):
        if gga:
            feats = np.stack([rho[0], (rho[1:4] ** 2).sum(0)])
            e, vr, vg = fwd(feats, None)
            return e, (vr, vg, None, None), None, None
        e, v = fwd(rho, None)
        return e, (v, None, None, None), None, None

    nelec, excsum, vmat = numint_ref.nr_rks(ao, weights, dm, eval_xc, xctype, hermi=hermi)
    D, tb = numint_ref.nr_rks_vjp(ao, weights, dm, fwd, vjp, e_bar, v_bar, xctype, hermi=hermi)
    return dict(nelec=nelec, excsum=excsum, vmat=vmat, dm_bar=D, theta_bar=tb)
