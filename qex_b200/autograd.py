"""Reverse-mode wiring of the hot path for a PyTorch host program.

The reference gets its VJP from ``jax.value_and_grad`` (trainer_legacy_no_jit.py:284); the shipped
JAX wiring is ``jax_ffi_shim.py`` (``jax.custom_vjp`` over the same two C entry points).  JAX is
not installable in this image, so this module exposes the identical forward/backward pair as a
``torch.autograd.Function`` -- it is what lets the end-to-end gradient (E_xc, V_xc) -> (dm, theta)
be exercised and tested here.  No arithmetic happens in Python.
"""
from __future__ import annotations

import torch

from .engine import XCContext


class _NrRks(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc: XCContext, xctype: str, hermi: int, dm: torch.Tensor, theta: torch.Tensor):
        out, resid = xc.nr_rks_fwd(dm, theta, xctype, hermi, want_resid=True)
        B, N = xc.nbatch, xc.nao
        ctx.xc, ctx.xctype, ctx.hermi = xc, xctype, hermi
        ctx.save_for_backward(theta.detach(), resid)
        vmat = out[:, : N * N].reshape(B, N, N).clone()
        excsum = out[:, N * N].clone()
        nelec = out[:, N * N + 1].clone()
        ctx.mark_non_differentiable(nelec)  # stop_grad in the reference (numint_legacy.py:305)
        return nelec, excsum, vmat

    @staticmethod
    def backward(ctx, _nelec_bar, e_bar, v_bar):
        theta, resid = ctx.saved_tensors
        xc = ctx.xc
        B, N = xc.nbatch, xc.nao
        e_bar = torch.zeros(B, dtype=torch.float64, device=theta.device) if e_bar is None else e_bar
        v_bar = torch.zeros(B, N, N, dtype=torch.float64, device=theta.device) if v_bar is None else v_bar
        bar = xc.nr_rks_vjp(theta, resid, e_bar.contiguous(), v_bar.contiguous(), ctx.xctype, ctx.hermi)
        return None, None, None, bar[: B * N * N].reshape(B, N, N), bar[B * N * N :]


def nr_rks(xc: XCContext, dm: torch.Tensor, theta: torch.Tensor, xctype: str = "NN", hermi: int = 0):
    """Differentiable ``(nelec [B], excsum [B], vmat [B,N,N])`` of the context's current grid/AO."""
    dm = dm.reshape(xc.nbatch, xc.nao, xc.nao)
    return _NrRks.apply(xc, xctype, hermi, dm, theta)


class _EvalRho(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc: XCContext, ncomp: int, hermi: int, dm: torch.Tensor):
        ctx.xc, ctx.ncomp, ctx.hermi = xc, ncomp, hermi
        return xc.eval_rho(dm, ncomp, hermi)

    @staticmethod
    def backward(ctx, rho_bar):
        return None, None, None, ctx.xc.eval_rho_vjp(rho_bar.contiguous(), ctx.ncomp, ctx.hermi)


def eval_rho(xc: XCContext, dm: torch.Tensor, ncomp: int = 1, hermi: int = 0):
    """Differentiable density on the grid (the trainer's density loss, trainer_legacy_no_jit.py:272-275)."""
    return _EvalRho.apply(xc, ncomp, hermi, dm.reshape(xc.nbatch, xc.nao, xc.nao))
