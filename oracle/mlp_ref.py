"""Oracle: the learned XC functional as an MLP (stage 3).  TEST INFRASTRUCTURE ONLY.

Restates, in NumPy float64:

* ``apply_local``  <- ``build_local_mlp.apply_fn`` qedft/models/classical/classical_models.py:168-172
                      (inputs / density_normalization_factor, vmap over points of
                      stax.serial(Dense, act, ..., Dense(1)), squeeze); layers from
                      ``build_mlp_layers`` :75-115, activations ``ACTIVATION_MAP`` :39-49.
* ``apply_global`` <- ``build_global_mlp.apply_fn`` classical_models.py:216-222.
* ``exc_and_vrho_local / _global`` <- qedft/train/td/trainer_legacy_no_jit.py:56-63 / :46-53
                      (``vrho := d(sum exc)/d rho``).
* ``*_vjp``        <- the second-order reverse rule JAX derives when the SCF energy is
                      differentiated w.r.t. the network parameters
                      (``trainer_legacy_no_jit.py:284``): cotangents of (exc, vrho) -> (rho, theta).
* flax ``MLP`` of trainer_legacy_no_jit.py:96-107 (gelu hidden layers, ``-scale*swish`` output) is
  covered by ``out_transform="neg_scale_swish"``.

stax ``Dense`` is ``x @ W + b`` with W [in, out]; ``stax.Gelu`` is the tanh-approximate GELU
(jax.nn.gelu default approximate=True).  parity unpinned against JAX itself (not installable);
derivatives pinned by torch float64 double-autograd and finite differences in tests/.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SELU_L = 1.0507009873554804934193349852946
SELU_A = 1.6732632423543772848170429916717
GELU_K = 0.7978845608028654  # sqrt(2/pi)
GELU_C = 0.044715

ACTIVATIONS = ("tanh", "relu", "softplus", "sigmoid", "elu", "leaky_relu", "selu", "gelu", "swish")


def act_d012(name, z):
    """sigma(z), sigma'(z), sigma''(z)."""
    if name == "tanh":
        t = np.tanh(z)
        d1 = 1.0 - t * t
        return t, d1, -2.0 * t * d1
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-z))
        d1 = s * (1.0 - s)
        return s, d1, d1 * (1.0 - 2.0 * s)
    if name == "softplus":
        s = 1.0 / (1.0 + np.exp(-z))
        return np.logaddexp(z, 0.0), s, s * (1.0 - s)
    if name == "relu":
        p = (z > 0).astype(z.dtype)
        return z * p, p, np.zeros_like(z)
    if name == "leaky_relu":
        p = np.where(z >= 0, 1.0, 0.01)
        return z * p, p, np.zeros_like(z)
    if name == "elu":
        e = np.exp(np.minimum(z, 0.0))
        pos = z > 0
        return np.where(pos, z, e - 1.0), np.where(pos, 1.0, e), np.where(pos, 0.0, e)
    if name == "selu":
        e = np.exp(np.minimum(z, 0.0))
        pos = z > 0
        return (
            SELU_L * np.where(pos, z, SELU_A * (e - 1.0)),
            SELU_L * np.where(pos, 1.0, SELU_A * e),
            SELU_L * np.where(pos, 0.0, SELU_A * e),
        )
    if name == "gelu":
        u = GELU_K * (z + GELU_C * z**3)
        u1 = GELU_K * (1.0 + 3.0 * GELU_C * z * z)
        u2 = GELU_K * 6.0 * GELU_C * z
        t = np.tanh(u)
        s = 1.0 - t * t
        f = 0.5 * z * (1.0 + t)
        d1 = 0.5 * (1.0 + t) + 0.5 * z * s * u1
        d2 = s * u1 + 0.5 * z * (-2.0 * t * s * u1 * u1 + s * u2)
        return f, d1, d2
    if name == "swish":
        s = 1.0 / (1.0 + np.exp(-z))
        ds = s * (1.0 - s)
        return z * s, s + z * ds, 2.0 * ds + z * ds * (1.0 - 2.0 * s)
    raise ValueError(f"Unknown activation '{name}'. Valid options: {list(ACTIVATIONS)}")


@dataclass
class MLPSpec:
    """sizes = [F, H1, ..., Hn, n_out]; hidden activation; input scale; output transform."""

    sizes: list
    activation: str = "tanh"
    in_scale: float = 0.5  # 1 / density_normalization_factor (classical_models.py:33,169)
    out_transform: str = "none"  # "none" | "neg_scale_swish" (flax MLP, trainer :107)
    out_scale: float = 1e-2
    n_dense: int = field(init=False)

    def __post_init__(self):
        self.n_dense = len(self.sizes) - 1

    def n_params(self):
        return sum(self.sizes[i] * self.sizes[i + 1] + self.sizes[i + 1] for i in range(self.n_dense))


def init_params(spec: MLPSpec, seed=0):
    """Glorot-normal W, N(0, 1e-2) b (stax.Dense defaults); seeded NumPy (not JAX's PRNG)."""
    rng = np.random.default_rng(seed)
    Ws, bs = [], []
    for i in range(spec.n_dense):
        fi, fo = spec.sizes[i], spec.sizes[i + 1]
        Ws.append(rng.standard_normal((fi, fo)) * np.sqrt(2.0 / (fi + fo)))
        bs.append(rng.standard_normal(fo) * 1e-2)
    return Ws, bs


def pack(Ws, bs):
    """flat theta = concat_l [W_l.ravel() (row-major [in,out]), b_l]."""
    return np.concatenate([np.concatenate([W.ravel(), b]) for W, b in zip(Ws, bs)])


def unpack(spec: MLPSpec, theta):
    Ws, bs, o = [], [], 0
    for i in range(spec.n_dense):
        fi, fo = spec.sizes[i], spec.sizes[i + 1]
        Ws.append(np.asarray(theta[o : o + fi * fo]).reshape(fi, fo))
        o += fi * fo
        bs.append(np.asarray(theta[o : o + fo]))
        o += fo
    return Ws, bs


def to_stax(Ws, bs):
    """stax.serial parameter list: [(W,b), (), (W,b), (), ..., (W,b)]."""
    out = []
    for i, (W, b) in enumerate(zip(Ws, bs)):
        out.append((W, b))
        if i < len(Ws) - 1:
            out.append(())
    return out


def from_stax(params):
    Ws = [p[0] for p in params if len(p) == 2]
    bs = [p[1] for p in params if len(p) == 2]
    return Ws, bs


# ----------------------------------------------------------------------------------------
# forward with tangent, reverse over both streams.  X is [P, F] (P samples of F features).
# ----------------------------------------------------------------------------------------
def _forward(spec, Ws, bs, X, Xdot=None):
    """Returns y [P, n_out], ydot (or None) and the tape."""
    h = X * spec.in_scale
    hd = None if Xdot is None else Xdot * spec.in_scale
    tape = []
    L = spec.n_dense
    for l in range(L):
        z = h @ Ws[l] + bs[l]
        zd = None if hd is None else hd @ Ws[l]
        last = l == L - 1
        if last and spec.out_transform == "none":
            tape.append((h, hd, z, zd, None, None, None))
            h, hd = z, zd
        else:
            name = spec.activation if not last else "swish"
            s0, s1, s2 = act_d012(name, z)
            c = 1.0 if not last else -spec.out_scale
            tape.append((h, hd, z, zd, c * s1, c * s2, name))
            h = c * s0
            hd = None if zd is None else c * s1 * zd
    return h, hd, tape


def apply(spec, Ws, bs, X):
    return _forward(spec, Ws, bs, X)[0]


def apply_local(spec, theta, inputs):
    """classical_models.py:168-172: inputs [G] or [G,F] -> [G]."""
    Ws, bs = unpack(spec, theta)
    X = np.asarray(inputs, dtype=np.float64)
    X = X.reshape(X.shape[0], -1)
    return apply(spec, Ws, bs, X).squeeze()


def apply_global(spec, theta, inputs):
    """classical_models.py:216-222: inputs [G] -> [n_out]."""
    Ws, bs = unpack(spec, theta)
    return apply(spec, Ws, bs, np.asarray(inputs, dtype=np.float64)[None, :])[0]


def _reverse(spec, Ws, tape, ybar, ydbar):
    """Reverse through (h, hdot) streams.  ybar/ydbar: cotangents of y / ydot [P, n_out].
    Returns Xbar [P,F], (Wbar list, bbar list)."""
    L = spec.n_dense
    hb, hdb = ybar, ydbar
    Wb, bb = [None] * L, [None] * L
    for l in reversed(range(L)):
        h, hd, z, zd, s1, s2, name = tape[l]
        if s1 is None:
            zb, zdb = hb, hdb
        else:
            zdb = None if hdb is None else hdb * s1
            zb = hb * s1
            if hdb is not None and zd is not None:
                zb = zb + hdb * s2 * zd
        Wb[l] = h.T @ zb
        if zdb is not None and hd is not None:
            Wb[l] = Wb[l] + hd.T @ zdb
        bb[l] = zb.sum(0)
        hb = zb @ Ws[l].T
        hdb = None if zdb is None else zdb @ Ws[l].T
    return hb * spec.in_scale, (Wb, bb)


def value_and_grad_x(spec, Ws, bs, X):
    """y = sum_out apply(X) per sample [P] and dy/dX [P,F] (one reverse pass)."""
    y, _, tape = _forward(spec, Ws, bs, X)
    xb, _ = _reverse(spec, Ws, tape, np.ones_like(y), None)
    return y.sum(1), xb


def exc_and_vrho_local(spec, theta, rho):
    """trainer_legacy_no_jit.py:56-63: exc [G], vrho [G] (per-point d exc_g / d rho_g)."""
    Ws, bs = unpack(spec, theta)
    exc, g = value_and_grad_x(spec, Ws, bs, np.asarray(rho, dtype=np.float64)[:, None])
    return exc, g[:, 0]


def exc_and_grad_features(spec, theta, feats):
    """GGA-feature extension: feats [F,G] -> exc [G], d exc/d feat [F,G]."""
    Ws, bs = unpack(spec, theta)
    exc, g = value_and_grad_x(spec, Ws, bs, np.asarray(feats, dtype=np.float64).T)
    return exc, g.T


def exc_and_vrho_global(spec, theta, rho):
    """trainer_legacy_no_jit.py:46-53: exc = sum(apply(rho)) scalar, vrho [G]."""
    Ws, bs = unpack(spec, theta)
    exc, g = value_and_grad_x(spec, Ws, bs, np.asarray(rho, dtype=np.float64)[None, :])
    return exc[0], g[0]


def second_order_vjp(spec, theta, X, ybar, gbar):
    """Cotangents (ybar [P], gbar [P,F]) of (y_p = sum_out apply(X_p), g_p = dy_p/dX_p) ->
    (Xbar [P,F], theta_bar flat).

    With L_p = ybar_p y_p + gbar_p . g_p:  gbar_p . g_p is the directional derivative of y_p
    along gbar_p, so run the forward pass with tangent Xdot = gbar, then reverse with seeds
    (ybar, 1) on (y, ydot).
    """
    Ws, bs = unpack(spec, theta)
    X = np.asarray(X, dtype=np.float64)
    y, yd, tape = _forward(spec, Ws, bs, X, Xdot=np.asarray(gbar, dtype=np.float64))
    seed = np.broadcast_to(np.asarray(ybar, dtype=np.float64)[:, None], y.shape)
    xb, (Wb, bb) = _reverse(spec, Ws, tape, seed, np.ones_like(y))
    return xb, pack(Wb, bb)


def exc_and_vrho_local_vjp(spec, theta, rho, exc_bar, vrho_bar):
    xb, tb = second_order_vjp(spec, theta, np.asarray(rho)[:, None], exc_bar, np.asarray(vrho_bar)[:, None])
    return xb[:, 0], tb


def exc_and_grad_features_vjp(spec, theta, feats, exc_bar, g_bar):
    xb, tb = second_order_vjp(spec, theta, np.asarray(feats).T, exc_bar, np.asarray(g_bar).T)
    return xb.T, tb


def exc_and_vrho_global_vjp(spec, theta, rho, exc_bar, vrho_bar):
    xb, tb = second_order_vjp(
        spec, theta, np.asarray(rho)[None, :], np.asarray([exc_bar], dtype=np.float64), np.asarray(vrho_bar)[None, :]
    )
    return xb[0], tb


def apply_vjp(spec, theta, X, ybar):
    """First-order VJP of ``apply_local`` (used by callers that differentiate apply_fn directly)."""
    Ws, bs = unpack(spec, theta)
    X = np.asarray(X, dtype=np.float64)
    X = X.reshape(X.shape[0], -1)
    y, _, tape = _forward(spec, Ws, bs, X)
    seed = np.broadcast_to(np.asarray(ybar, dtype=np.float64).reshape(-1, 1), y.shape)
    xb, (Wb, bb) = _reverse(spec, Ws, tape, seed, None)
    return xb, pack(Wb, bb)
