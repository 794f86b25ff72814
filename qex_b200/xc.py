"""Host mirror of the reference's XC-functional glue (stage 3).

``eval_xc`` keeps the signature of ``qedft/train/td/trainer_legacy_no_jit.py:76-93``;
``get_eval_xc`` that of ``qedft/train/td/xc.py:57-87`` (raises ``ValueError`` when ``deriv != 1``,
:73-74).  ``exc_and_vrho_local / _global`` are :56-63 / :46-53: ``vrho := d(sum exc)/d rho``.
The numbers come from ``qexxc_xc_fwd`` (CUDA); a ``network`` whose ``apply`` is not one of this
package's native functionals is rejected -- there is no host evaluation path here.
"""
from __future__ import annotations

from functools import partial

import numpy as np

from . import _lib
from .engine import XCContext


def _native_apply(network):
    """Accept (init_fn, apply_fn), a StaxAdapter-like object with .apply_fn, or the apply_fn itself."""
    if isinstance(network, tuple) and len(network) == 2:
        network = network[1]
    fn = getattr(network, "apply_fn", network)
    if not hasattr(fn, "qex_spec"):
        raise TypeError(
            "network is not a qex_b200 native functional (build it with qex_b200.networks); arbitrary Python "
            "functionals go through NumInt.eval_xc as a host callback instead")
    return fn


def _theta(fn, params):
    if isinstance(params, dict) and "params" in params:  # flax-style wrapper of StaxAdapter
        params = params["params"]["stax_params"]
    return fn.flatten(params)


def _ctx_for(fn, ngrids, ncomp=1) -> XCContext:
    ctx = fn.native.ctx(ngrids)
    if ctx.ncomp < ncomp:
        raise ValueError("this functional was built without gradient (GGA) features")
    return ctx


def _out(t, like):
    import torch

    return t if isinstance(like, torch.Tensor) else t.cpu().numpy()


def exc_and_vrho_local(network, params, rho):
    """trainer_legacy_no_jit.py:56-63."""
    fn = _native_apply(network)
    G = int(np.prod(rho.shape))
    ctx = _ctx_for(fn, G)
    ctx.set_grid(None, np.ones(G)) if ctx.ngrids != G else None
    exc, vrho, _ = ctx.xc_fwd(rho, _theta(fn, params), "NN")
    return _out(exc[0], rho), _out(vrho[0], rho)


def exc_and_vrho_global(network, params, rho):
    """trainer_legacy_no_jit.py:46-53: exc is the scalar sum of the network output."""
    fn = _native_apply(network)
    if fn.qex_spec.kind != _lib.NET_GLOBAL_MLP:
        # a per-point network under the global flag (the reference's default `is_global_xc=True`):
        # exc = jnp.sum(network.apply(params, rho)), vrho = its per-point derivative
        exc, vrho = exc_and_vrho_local(network, params, rho)
        return exc.sum(), vrho
    G = int(np.prod(rho.shape))
    ctx = _ctx_for(fn, G)
    ctx.set_grid(None, np.ones(G)) if ctx.ngrids != G else None
    exc, vrho, _ = ctx.xc_fwd(rho, _theta(fn, params), "NN-AmplitudeEncoding")
    return _out(exc[0], rho), _out(vrho[0], rho)


def eval_xc(xc_code, rho, spin=0, relativity=0, deriv=2, verbose=None, params=None, network=None,
            is_global_xc=True):
    """trainer_legacy_no_jit.py:76-93 -> (exc, (vrho, None, None, None), None, None)."""
    if is_global_xc:
        exc, vrho = exc_and_vrho_global(network, params, rho)
    else:
        exc, vrho = exc_and_vrho_local(network, params, rho)
    return exc, (vrho, None, None, None), None, None


LDA_EXCHANGE_CODES = ("lda", "lda,", "slater", "slater,", "lda_x", "lda_x,")  # pyscf aliases of libxc LDA_X alone


def lda_exchange(rho):
    """libxc LDA_X, unpolarised, on the device (csrc/xc_lda.cu): rho (CUDA tensor or array, any shape) ->
    (exc per particle, vrho = d(rho exc)/d rho), returned in the type that came in."""
    import ctypes as C

    import torch

    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.QexxcError(_lib.ERR_NODEVICE, "no CUDA device: qex_b200 has no CPU fallback")
    if isinstance(rho, torch.Tensor) and rho.is_cuda:
        r = rho.to(torch.float64).contiguous()
    else:
        r = torch.as_tensor(np.ascontiguousarray(rho, dtype=np.float64)).cuda()
    exc, vrho = torch.empty_like(r), torch.empty_like(r)
    with torch.cuda.device(r.device):
        _lib.check(lib.qexxc_lda_exchange(r.device.index, r.data_ptr(), r.numel(), exc.data_ptr(), vrho.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return _out(exc, rho), _out(vrho, rho)


def lda_eval_xc(xc_code, rho, spin=0, relativity=0, deriv=1, verbose=None, params=None, **kwargs):
    """``ni.eval_xc("lda", rho, ...)`` of the reference's LDA branch (pyscfad libxc) for Slater exchange:
    -> (exc, (vrho, None, None, None), None, None).  Other libxc functionals stay with pyscf."""
    if not (isinstance(xc_code, str) and xc_code.lower().replace(" ", "") in LDA_EXCHANGE_CODES):
        raise NotImplementedError(f"xc_code {xc_code!r}: only Slater exchange ('lda') has a kernel; other libxc "
                                  "functionals stay with pyscf")
    if spin != 0:
        raise NotImplementedError("spin-polarised LDA is not on the accelerated path")
    exc, vrho = lda_exchange(rho)
    return exc, (vrho, None, None, None), None, None


def make_eval_xc(network, is_global_xc=False):
    """What the trainer installs with ``mf.define_xc_(description=...)``
    (trainer_legacy_no_jit.py:256-261): ``eval_xc`` with the network bound."""
    _native_apply(network)
    return partial(eval_xc, network=network, is_global_xc=is_global_xc)


def get_eval_xc(xc_code, rho, spin=0, relativity=0, deriv=0, verbose=None, params=None, network=None, **kwargs):
    """qedft/train/td/xc.py:57-87 (rho is the 4-tuple (rho0, dx, dy, dz); only rho0 is used)."""
    if deriv != 1:
        raise ValueError("eval_xc: deriv should be set to 1.")
    rho0 = rho[0]
    exc, vrho = exc_and_vrho_local(network, params, rho0)
    return exc, (vrho, None, None, None), None, None
