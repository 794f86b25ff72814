"""jax.ffi registration + jax.custom_vjp wrappers over libqexxc.so.

UNTESTED HERE: JAX is not installable in this repository's environment (no network, not in the
wheelhouse), so this module is import-guarded source showing exactly what a maintainer of the
reference adds to make ``nr_rks`` / ``eval_rho`` / ``apply_fn`` custom calls
(INTEGRATION.md).  Everything it calls on the C side is exercised by the test
suite through ctypes; qex_b200/autograd.py is the same forward/backward pair wired into torch.

Usage in the reference (qedft/train/td/numint_legacy.py:588)::

    from qex_b200 import jax_ffi_shim as qx
    qx.register()
    handle = qx.Handle(ctx)                      # an engine.XCContext with grid + AO loaded
    NumInt.nr_rks = qx.make_nr_rks(handle, xctype="NN")   # same (nelec, excsum, vmat) return
"""
from __future__ import annotations

import subprocess
from pathlib import Path

import numpy as np

try:  # pragma: no cover - JAX is absent in this image
    import jax
    import jax.numpy as jnp

    HAVE_JAX = True
except Exception:  # noqa: BLE001
    jax = jnp = None
    HAVE_JAX = False

HERE = Path(__file__).resolve().parent
_TARGETS = {
    "qexxc_nr_rks_fwd": "QexxcNrRksFwd", "qexxc_nr_rks_vjp": "QexxcNrRksVjp", "qexxc_eval_rho": "QexxcEvalRho",
    "qexxc_eval_rho_vjp": "QexxcEvalRhoVjp", "qexxc_apply_fn_fwd": "QexxcApplyFwd", "qexxc_apply_fn_vjp": "QexxcApplyVjp",
}


def _require():
    if not HAVE_JAX:
        raise ImportError("jax is not installed: qex_b200.jax_ffi_shim needs jax[cuda12] >= 0.4.35 (jax.ffi)")


def build_handlers() -> Path:  # pragma: no cover
    """g++ the XLA-FFI adaptor (csrc/jax_ffi/qexxc_jax_ffi.cc) against jaxlib's headers."""
    _require()
    out = HERE / "libqexxc_jax.so"
    src = HERE / "csrc" / "jax_ffi" / "qexxc_jax_ffi.cc"
    if not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call([
            "g++", "-shared", "-fPIC", "-std=c++17", "-O2", f"-I{jax.ffi.include_dir()}", f"-I{HERE.parent / 'include'}",
            str(src), f"-L{HERE}", "-lqexxc", f"-Wl,-rpath,{HERE}", "-o", str(out)])
    return out


def register():  # pragma: no cover
    """jax.ffi.register_ffi_target for every handler, platform CUDA."""
    _require()
    import ctypes

    lib = ctypes.CDLL(str(build_handlers()))
    for target, sym in _TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, sym)), platform="CUDA")


class Handle:
    """Static description of a loaded engine.XCContext that the traced functions close over."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.addr = np.int64(ctx._h.value)
        self.B, self.N, self.G = ctx.nbatch, ctx.nao, ctx.ngrids
        self.n_out = self.N * self.N + 2
        self.n_resid = ctx.resid_doubles


def make_nr_rks(h: Handle, xctype: str = "NN", hermi: int = 0):  # pragma: no cover
    """-> f(dm, theta) = (nelec, excsum, vmat) with a custom VJP (nelec is stop-gradient,
    numint_legacy.py:305).  Drop-in body for NumInt.nr_rks once mol/grids are bound to `h`."""
    _require()
    from .engine import XCTYPES

    xt = np.int32(XCTYPES[xctype])
    hm = np.int32(hermi)
    f64 = jnp.float64

    def fwd_call(dm, theta):
        out, resid = jax.ffi.ffi_call(
            "qexxc_nr_rks_fwd",
            (jax.ShapeDtypeStruct((h.B, h.n_out), f64), jax.ShapeDtypeStruct((h.n_resid,), f64)),
        )(dm.reshape(h.B, h.N, h.N), theta, ctx=h.addr, xctype=xt, hermi=hm)
        return out, resid

    def unpack(out):
        vmat = out[:, : h.N * h.N].reshape(h.B, h.N, h.N)
        return jax.lax.stop_gradient(out[:, h.N * h.N + 1]), out[:, h.N * h.N], vmat

    @jax.custom_vjp
    def nr_rks(dm, theta):
        return unpack(fwd_call(dm, theta)[0])

    def nr_rks_fwd(dm, theta):
        out, resid = fwd_call(dm, theta)
        return unpack(out), (theta, resid)

    def nr_rks_bwd(res, cot):
        theta, resid = res
        _, e_bar, v_bar = cot
        bar = jax.ffi.ffi_call("qexxc_nr_rks_vjp", jax.ShapeDtypeStruct((h.B * h.N * h.N + theta.shape[0],), f64))(
            theta, resid, e_bar.reshape(h.B), v_bar.reshape(h.B, h.N, h.N), ctx=h.addr, xctype=xt, hermi=hm)
        return bar[: h.B * h.N * h.N].reshape(h.B, h.N, h.N), bar[h.B * h.N * h.N :]

    nr_rks.defvjp(nr_rks_fwd, nr_rks_bwd)
    return nr_rks


def make_eval_rho(h: Handle, ncomp: int = 1, hermi: int = 0):  # pragma: no cover
    """-> f(dm) = rho [B, ncomp, G] with a custom VJP (numint_legacy.py:351-397)."""
    _require()
    nc, hm, f64 = np.int32(ncomp), np.int32(hermi), jnp.float64

    @jax.custom_vjp
    def eval_rho(dm):
        return jax.ffi.ffi_call("qexxc_eval_rho", jax.ShapeDtypeStruct((h.B, ncomp, h.G), f64))(
            dm.reshape(h.B, h.N, h.N), ctx=h.addr, ncomp=nc, hermi=hm)

    def fwd(dm):
        return eval_rho(dm), None

    def bwd(_, rho_bar):
        return (jax.ffi.ffi_call("qexxc_eval_rho_vjp", jax.ShapeDtypeStruct((h.B, h.N, h.N), f64))(
            rho_bar, ctx=h.addr, ncomp=nc, hermi=hm),)

    eval_rho.defvjp(fwd, bwd)
    return eval_rho


def make_apply_fn(h: Handle, n_theta: int):  # pragma: no cover
    """-> apply_fn(theta, inputs) -> [npts] with a custom VJP (networks.py:43-75)."""
    _require()
    f64 = jnp.float64

    @jax.custom_vjp
    def apply_fn(theta, x):
        n = np.int64(x.shape[0])
        return jax.ffi.ffi_call("qexxc_apply_fn_fwd", jax.ShapeDtypeStruct((x.shape[0],), f64))(
            x, theta, ctx=h.addr, npts=n)

    def fwd(theta, x):
        return apply_fn(theta, x), (theta, x)

    def bwd(res, y_bar):
        theta, x = res
        xb, tb = jax.ffi.ffi_call(
            "qexxc_apply_fn_vjp", (jax.ShapeDtypeStruct(x.shape, f64), jax.ShapeDtypeStruct((n_theta,), f64)))(
            x, theta, y_bar, ctx=h.addr, npts=np.int64(x.shape[0]))
        return tb, xb

    apply_fn.defvjp(fwd, bwd)
    return apply_fn
