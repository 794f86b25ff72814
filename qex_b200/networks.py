"""Host mirror of the reference's network interface for the per-grid-point functionals.

Same names, argument meaning and error behaviour as ``qedft/models/networks.py``:
``LocalMLP`` (:83-109), ``GlobalMLP`` (:112-138), ``LocalQNN`` (:180-228), each with
``build_network(grids) -> (init_fn, apply_fn)``; ``init_fn(rng, input_shape) -> (output_shape,
params)`` and ``apply_fn(params, inputs, **kwargs) -> outputs`` as in
``classical_models.py:163-174`` / ``quantum_models.py:740-776``.  Parameters keep the reference's
structure (stax list ``[(W, b), (), (W, b), ...]`` for the MLPs, a flat vector for the QNN).

``apply_fn`` runs the CUDA kernels through the C ABI (``qexxc_apply_fn_fwd``); it also carries
``.qex_spec`` / ``.flatten`` / ``.unflatten`` so that ``numint.nr_rks`` and ``xc.eval_xc`` can take
the fused device path.  There is no host implementation of the forward pass here.
"""
from __future__ import annotations

from collections.abc import Callable

import numpy as np

from . import _lib
from .engine import NetSpec, XCContext

DEFAULT_N_NEURONS = 64
DEFAULT_N_LAYERS = 3
DEFAULT_ACTIVATION = "tanh"
DEFAULT_DENSITY_NORM = 2.0
# classical_models.py:39-49 (softmax is not element-wise and has no kernel here)
ACTIVATION_MAP = ("tanh", "relu", "softplus", "sigmoid", "elu", "leaky_relu", "selu", "gelu")


def _rng(rng):
    """int seed, numpy Generator, or a JAX PRNGKey-like integer array."""
    if isinstance(rng, np.random.Generator):
        return rng
    arr = np.asarray(rng).astype(np.uint64).ravel()
    return np.random.default_rng([int(v) for v in arr] if arr.size > 1 else int(arr[0]))


def _num_grids(grids):
    if grids is None:
        raise ValueError("grids must be provided")
    g = getattr(grids, "coords", grids)
    return int(np.asarray(g).shape[0]) if not hasattr(g, "shape") else int(g.shape[0])


# ---- stax-style parameter structure <-> flat theta ------------------------------------------------
def stax_flatten(params) -> np.ndarray:
    """[(W [in,out], b [out]), (), ...] -> concat_l [W_l.ravel(), b_l] (float64)."""
    parts = []
    for p in params:
        if len(p) == 2:
            parts.append(np.asarray(p[0], dtype=np.float64).ravel())
            parts.append(np.asarray(p[1], dtype=np.float64).ravel())
    return np.concatenate(parts)


def stax_unflatten(theta, sizes):
    theta = np.asarray(theta, dtype=np.float64)
    out, o = [], 0
    for i, (fi, fo) in enumerate(zip(sizes[:-1], sizes[1:])):
        W = theta[o : o + fi * fo].reshape(fi, fo)
        o += fi * fo
        b = theta[o : o + fo]
        o += fo
        out.append((W, b))
        if i < len(sizes) - 2:
            out.append(())
    return out


def _stax_init(rng, sizes):
    """stax.Dense defaults: glorot-normal W, N(0, 1e-2) b (seeded numpy, not JAX's PRNG stream)."""
    g = _rng(rng)
    params = []
    for i, (fi, fo) in enumerate(zip(sizes[:-1], sizes[1:])):
        W = g.standard_normal((fi, fo)) * np.sqrt(2.0 / (fi + fo))
        b = g.standard_normal(fo) * 1e-2
        params.append((W, b))
        if i < len(sizes) - 2:
            params.append(())
    return params


class _Native:
    """Lazily created device context shared by the apply_fn closures of one network."""

    def __init__(self, spec: NetSpec, capacity: int):
        self.spec, self.capacity, self._ctx = spec, int(capacity), None

    def ctx(self, npts: int) -> XCContext:
        if self._ctx is None or npts > self._ctx.ngrids_max:
            if self._ctx is not None:
                self._ctx.close()
            ncomp = 4 if (self.spec.kind == _lib.NET_LOCAL_MLP and self.spec.n_features > 1) else 1
            self._ctx = XCContext(nao=1, ngrids_max=max(npts, self.capacity), ncomp=ncomp, net=self.spec)
        return self._ctx


def _finish(y, like):
    import torch

    if isinstance(like, torch.Tensor):
        return y
    return y.cpu().numpy()


def _validate_mlp(n_neurons, n_layers, activation):
    if n_neurons <= 0:
        raise ValueError("n_neurons must be positive")
    if n_layers <= 0:
        raise ValueError("n_layers must be positive")
    if activation not in ACTIVATION_MAP:
        raise ValueError(f"Unknown activation '{activation}'. Valid options: {list(ACTIVATION_MAP)}")


def build_local_mlp(n_neurons=DEFAULT_N_NEURONS, n_layers=DEFAULT_N_LAYERS, activation=DEFAULT_ACTIVATION,
                    n_outputs=1, density_normalization_factor=DEFAULT_DENSITY_NORM, grids=None, n_features=1,
                    precision="f64", **kwargs) -> tuple[Callable, Callable]:
    """classical_models.py:123-174.  ``n_features`` > 1 is the GGA-feature extension."""
    if grids is None:
        raise ValueError("grids must be provided for local MLP")
    _validate_mlp(n_neurons, n_layers, activation)
    if n_outputs != 1:
        raise NotImplementedError("local MLP kernels produce one output per grid point")
    num_grids = _num_grids(grids)
    sizes = [n_features] + [n_neurons] * n_layers + [1]
    spec = NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=n_features, n_hidden=n_layers, width=n_neurons,
                   activation=activation, precision=precision, in_scale=1.0 / density_normalization_factor)
    native = _Native(spec, num_grids)

    def init_fn(rng, input_shape):
        del input_shape
        return (-1, num_grids, 1), _stax_init(rng, sizes)

    def apply_fn(params, inputs, **kwargs):
        del kwargs
        theta = stax_flatten(params)
        npts = int(np.prod(inputs.shape)) // n_features
        y = native.ctx(npts).apply_fn(inputs, theta)
        return _finish(y, inputs)

    apply_fn.qex_spec = spec
    apply_fn.flatten = stax_flatten
    apply_fn.unflatten = lambda theta: stax_unflatten(theta, sizes)
    apply_fn.native = native
    return init_fn, apply_fn


def build_global_mlp(n_neurons=DEFAULT_N_NEURONS, n_layers=DEFAULT_N_LAYERS, activation=DEFAULT_ACTIVATION,
                     n_outputs=1, density_normalization_factor=DEFAULT_DENSITY_NORM, grids=None,
                     **kwargs) -> tuple[Callable, Callable]:
    """classical_models.py:177-224: the network sees the whole density vector."""
    if grids is None:
        raise ValueError("grids must be provided for global MLP")
    _validate_mlp(n_neurons, n_layers, activation)
    if n_outputs != 1:
        raise NotImplementedError("global MLP kernel produces one output")
    num_grids = _num_grids(grids)
    sizes = [num_grids] + [n_neurons] * n_layers + [1]
    spec = NetSpec(kind=_lib.NET_GLOBAL_MLP, n_hidden=n_layers, width=n_neurons, activation=activation,
                   in_scale=1.0 / density_normalization_factor)
    native = _Native(spec, num_grids)

    def init_fn(rng, input_shape):
        del input_shape
        return (1,), _stax_init(rng, sizes)

    def apply_fn(params, inputs, **kwargs):
        del kwargs
        if int(np.prod(inputs.shape)) != num_grids:
            raise ValueError(f"global MLP was built for {num_grids} grid points, got {tuple(inputs.shape)}")
        y = native.ctx(num_grids).apply_fn(inputs, stax_flatten(params))
        return _finish(y, inputs)

    apply_fn.qex_spec = spec
    apply_fn.flatten = stax_flatten
    apply_fn.unflatten = lambda theta: stax_unflatten(theta, sizes)
    apply_fn.native = native
    return init_fn, apply_fn


def build_qnn(n_qubits: int, n_features: int = 1, n_layers: int = 2, qnn_type: str = "LocalQNN",
              layer_type: str = "DirectQNN", grids=None, precision="f64", **kwargs) -> tuple[Callable, Callable]:
    """quantum_models.py:623-776 for the per-grid-point case: DirectQNN feature map + hea(n, L)."""
    if qnn_type != "LocalQNN":
        raise ValueError(f"Unsupported QNN type: {qnn_type}. Only 'LocalQNN' is on the accelerated path.")
    if layer_type != "DirectQNN":
        raise NotImplementedError(f"layer_type {layer_type!r}: only the DirectQNN feature map has a kernel")
    if kwargs.get("noise") is not None or kwargs.get("n_shots", 0):
        raise NotImplementedError("noisy / shot-based simulation is not on the accelerated path")
    num_grids = _num_grids(grids)
    n_params = 3 * n_qubits * n_layers
    spec = NetSpec(kind=_lib.NET_LOCAL_QNN, n_features=1, n_hidden=n_layers, width=n_qubits, precision=precision,
                   in_scale=1.0)
    native = _Native(spec, num_grids)

    def init_fn(rng, input_shape):
        del input_shape
        return (-1, num_grids, 1), _rng(rng).uniform(-0.1, 0.1, n_params)  # quantum_models.py:752-757

    def apply_fn(params, inputs, **kwargs):
        del kwargs
        shp = tuple(inputs.shape)
        # LocalQNN.__call__ rejects 2-D input (quantum_models.py:487-492); the 3D local path passes
        # x[:, None], so [G, 1] is accepted here and squeezed
        if len(shp) > 2 or (len(shp) == 2 and shp[1] != 1):
            raise ValueError(f"LocalQNN expects inputs of shape (N,) for local processing, but got shape {shp}.")
        y = native.ctx(int(np.prod(shp))).apply_fn(inputs, np.asarray(params, dtype=np.float64).ravel())
        return _finish(y, inputs)

    apply_fn.qex_spec = spec
    apply_fn.flatten = lambda p: np.asarray(p, dtype=np.float64).ravel()
    apply_fn.unflatten = lambda theta: np.asarray(theta, dtype=np.float64)
    apply_fn.native = native
    return init_fn, apply_fn


# ---- the class zoo (networks.py) -------------------------------------------------------------------
class KohnShamNetwork:
    """networks.py:43-75."""

    def __init__(self, config_dict: dict | None = None):
        self.config = {}
        if config_dict is not None:
            self.config.update(config_dict)

    def build_network(self, grids):
        raise NotImplementedError


class LocalMLP(KohnShamNetwork):
    """networks.py:83-109."""

    def __init__(self, config_dict: dict | None = None):
        self.config = {"network_type": "mlp", "wrap_self_interaction": False, "wrap_with_negative_transform": True,
                       "use_amplitude_encoding": False}
        if config_dict is not None:
            self.config.update(config_dict)
        if self.config.get("use_amplitude_encoding") is True:
            raise ValueError("Set use_amplitude_encoding to False.")

    def build_network(self, grids):
        return build_local_mlp(
            n_neurons=self.config.get("n_neurons", 64), n_layers=self.config.get("n_layers", 3),
            activation=self.config.get("activation", "tanh"),
            density_normalization_factor=self.config.get("density_normalization_factor", 2.0), grids=grids,
            n_features=self.config.get("n_features", 1), precision=self.config.get("precision", "f64"))


class GlobalMLP(KohnShamNetwork):
    """networks.py:112-138."""

    def __init__(self, config_dict: dict | None = None):
        self.config = {"network_type": "mlp_ksr", "wrap_self_interaction": True, "wrap_with_negative_transform": True,
                       "use_amplitude_encoding": True}
        if config_dict is not None:
            self.config.update(config_dict)
        if self.config.get("use_amplitude_encoding") is False:
            raise ValueError("Set use_amplitude_encoding to True.")

    def build_network(self, grids):
        return build_global_mlp(
            n_neurons=self.config.get("n_neurons", 64), n_layers=self.config.get("n_layers", 3),
            activation=self.config.get("activation", "tanh"),
            density_normalization_factor=self.config.get("density_normalization_factor", 2.0), grids=grids)


class LocalQNN(KohnShamNetwork):
    """networks.py:180-228 (defaults: n_qubits 2, n_layers 2, DirectQNN, zero initial state)."""

    def __init__(self, config_dict: dict | None = None, noise=None):
        self.config = {"network_type": "mlp", "wrap_self_interaction": False, "wrap_with_negative_transform": True,
                       "use_amplitude_encoding": False, "qnn_type": "LocalQNN", "layer_type": "DirectQNN",
                       "map_fn": None}
        if config_dict is not None:
            self.config.update(config_dict)
        self.noise = noise

    def build_network(self, grids, noise=None):
        if self.noise is None:
            self.noise = noise
        return build_qnn(
            n_qubits=self.config.get("n_qubits", 2), n_layers=self.config.get("n_layers", 2),
            qnn_type=self.config.get("qnn_type", "LocalQNN"), layer_type=self.config.get("layer_type", "DirectQNN"),
            grids=grids, n_features=self.config.get("n_features", 1), noise=self.noise,
            n_shots=self.config.get("n_shots", 0), precision=self.config.get("precision", "f64"))


class StaxAdapter:
    """qedft/train/td/stax_to_flax_network.py:9-46: flax-style ``apply({"params": {"stax_params": p}}, x)``."""

    def __init__(self, init_fn, apply_fn, input_shape, rng_key=0):
        self.init_fn, self.apply_fn = init_fn, apply_fn
        _, self.stax_params = init_fn(rng_key, (-1,) + tuple(input_shape))

    def init(self, rng_key, inputs):
        return {"params": {"stax_params": self.stax_params}}

    def apply(self, params, inputs, **kwargs):
        stax_params = params["params"]["stax_params"] if isinstance(params, dict) and "params" in params else params
        return self.apply_fn(stax_params, inputs)


def adapt_stax_for_training(init_fn, apply_fn, input_shape, rng_key=0):
    """stax_to_flax_network.py:48-65."""
    adapter = StaxAdapter(init_fn, apply_fn, input_shape, rng_key)
    return adapter, adapter.init(rng_key, None)
