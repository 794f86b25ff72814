"""Host mirror of the reference's incore Coulomb / exchange build.

Same names and argument meaning as qedft/train/td/hf_legacy.py:
    `_dot_eri_dm_s1(eri, dm, with_j, with_k)`  :275-286
    `dot_eri_dm(eri, dm, hermi, with_j, with_k)` :289-299
    `make_rdm1(mo_coeff, mo_occ)` :331-338      (dense N x N product, torch)
and of the Coulomb energy line of `get_veff` (rks_legacy.py:122).  Every J/K number comes from
the kernels in csrc/jk.cu through the C ABI (qexxc_dot_eri_dm / qexxc_dot_eri_dm_vjp); there is
no CPU or torch fallback.  The ERI tensor should be passed as a float64 CUDA tensor so that it
stays resident in HBM across SCF cycles (a numpy array is uploaded on every call).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

_work: dict = {}


def _dev(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def _device_of(eri) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.QexxcError(_lib.ERR_NODEVICE, "no CUDA device: qex_b200 has no CPU fallback")
    if isinstance(eri, torch.Tensor) and eri.is_cuda:
        return eri.device
    return torch.device("cuda", torch.cuda.current_device())


def _workspace(lib, device: torch.device, nao: int, nmol: int = 1) -> torch.Tensor:
    key = (device.index, nao, nmol)
    if key not in _work:
        n = C.c_long()
        check(lib.qexxc_jk_workspace_doubles(device.index, nao, C.byref(n)))
        _work[key] = torch.empty(max(n.value * nmol, 1), dtype=torch.float64, device=device)
    return _work[key]


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dot_eri_dm_s1(eri, dm, with_j=True, with_k=True):
    """vj = einsum("ijkl,xji->xkl"), vk = einsum("ijkl,xjk->xil") in one pass over `eri`.
    Returns CUDA tensors shaped like `dm` (None for a matrix that was not requested)."""
    lib = _lib.load()
    device = _device_of(eri)
    nao = int(dm.shape[-1])
    e = _dev(eri, device)
    if e.numel() != nao**4:
        raise ValueError(f"eri has {e.numel()} elements, nao**4 = {nao**4} expected")
    d = _dev(dm, device)
    nset = d.numel() // (nao * nao)
    vj = torch.empty_like(d) if with_j else None
    vk = torch.empty_like(d) if with_k else None
    w = _workspace(lib, device, nao)
    with torch.cuda.device(device):
        check(lib.qexxc_dot_eri_dm(device.index, e.data_ptr(), d.data_ptr(), nset, nao, int(bool(with_j)),
                                   int(bool(with_k)), vj.data_ptr() if with_j else None,
                                   vk.data_ptr() if with_k else None, w.data_ptr(), w.numel(), _stream(device)))
    return vj, vk


def _dot_eri_dm_s1_vjp(eri, nao, vj_bar=None, vk_bar=None):
    """dm cotangent of `_dot_eri_dm_s1` (the transposed contractions), one pass over `eri`."""
    lib = _lib.load()
    device = _device_of(eri)
    e = _dev(eri, device)
    ref = vj_bar if vj_bar is not None else vk_bar
    if ref is None:
        raise ValueError("at least one cotangent is needed")
    a = _dev(vj_bar, device) if vj_bar is not None else None
    b = _dev(vk_bar, device) if vk_bar is not None else None
    shape = tuple(ref.shape)
    nset = int(np.prod(shape)) // (nao * nao)
    out = torch.empty(shape, dtype=torch.float64, device=device)
    w = _workspace(lib, device, nao)
    with torch.cuda.device(device):
        check(lib.qexxc_dot_eri_dm_vjp(device.index, e.data_ptr(), a.data_ptr() if a is not None else None,
                                       b.data_ptr() if b is not None else None, nset, nao, out.data_ptr(),
                                       w.data_ptr(), w.numel(), _stream(device)))
    return out


def dot_eri_dm(eri, dm, hermi=0, with_j=True, with_k=True):
    """hf_legacy.py:289-299.  `hermi` is accepted and unused, as in the reference's dense branch."""
    nao = int(dm.shape[-1])
    size = eri.numel() if isinstance(eri, torch.Tensor) else np.size(eri)
    if eri.is_complex() if isinstance(eri, torch.Tensor) else np.iscomplexobj(eri):
        raise NotImplementedError("complex ERI")
    if size != nao**4:
        # the reference hands packed s4/s8 tensors to pyscfad's _vhf.incore; not on the accelerated path
        raise NotImplementedError("only the dense s1 tensor (eri.size == nao**4) is on the accelerated path")
    return _dot_eri_dm_s1(eri, dm, with_j, with_k)


class _DotEriDm(torch.autograd.Function):
    """torch.autograd bridge: gradient w.r.t. dm only (the tensor is data, as in the reference)."""

    @staticmethod
    def forward(ctx, eri, dm, with_j, with_k):
        vj, vk = _dot_eri_dm_s1(eri, dm, with_j, with_k)
        ctx.eri, ctx.nao, ctx.wj, ctx.wk = eri, int(dm.shape[-1]), with_j, with_k
        z = dm.new_zeros(())
        return (vj if with_j else z), (vk if with_k else z)

    @staticmethod
    def backward(ctx, vj_bar, vk_bar):
        a = vj_bar.contiguous() if ctx.wj else None
        b = vk_bar.contiguous() if ctx.wk else None
        return None, _dot_eri_dm_s1_vjp(ctx.eri, ctx.nao, a, b), None, None


def dot_eri_dm_autograd(eri, dm, with_j=True, with_k=True):
    vj, vk = _DotEriDm.apply(eri, dm, with_j, with_k)
    return (vj if with_j else None), (vk if with_k else None)


class _RowDot(torch.autograd.Function):
    """J[i,j] = sum_kl eri[i,j,k,l] dm[k,l] (the `einsum("ijkl,kl->ij")` of scf_functions_masked.py:152):
    forward is the J-bar half of the reverse kernel read back transposed, backward is the forward kernel."""

    @staticmethod
    def forward(ctx, eri, dm):
        ctx.eri = eri
        nao = int(dm.shape[-1])
        return _dot_eri_dm_s1_vjp(eri, nao, dm.contiguous(), None).transpose(-1, -2).contiguous()

    @staticmethod
    def backward(ctx, j_bar):
        vj, _ = _dot_eri_dm_s1(ctx.eri, j_bar.transpose(-1, -2).contiguous(), True, False)
        return None, vj


def dot_eri_dm_rowdot(eri, dm):
    return _RowDot.apply(eri, dm)


# ---- batched over molecules of equal nao (the c4 pattern: one launch for a whole dissociation curve) -------
def dot_eri_dm_batched(eri, dm, with_j=True, with_k=True):
    """eri [B, nao^4 ...], dm [B, nao, nao] -> (vj, vk) [B, nao, nao] with the einsums of `_dot_eri_dm_s1`
    applied per molecule, all molecules in one launch."""
    lib = _lib.load()
    device = _device_of(eri)
    nao, nmol = int(dm.shape[-1]), int(dm.shape[0])
    e, d = _dev(eri, device), _dev(dm, device)
    if e.numel() != nmol * nao**4 or d.numel() != nmol * nao * nao:
        raise ValueError("eri must be [B, nao^4] and dm [B, nao, nao]")
    vj = torch.empty_like(d) if with_j else None
    vk = torch.empty_like(d) if with_k else None
    w = _workspace(lib, device, nao, nmol)
    with torch.cuda.device(device):
        check(lib.qexxc_dot_eri_dm_batched(device.index, e.data_ptr(), d.data_ptr(), nmol, nao, int(bool(with_j)),
                                           int(bool(with_k)), vj.data_ptr() if with_j else None,
                                           vk.data_ptr() if with_k else None, w.data_ptr(), w.numel(), _stream(device)))
    return vj, vk


def _dot_eri_dm_batched_vjp(eri, nao, nmol, vj_bar=None, vk_bar=None):
    lib = _lib.load()
    device = _device_of(eri)
    e = _dev(eri, device)
    a = _dev(vj_bar, device) if vj_bar is not None else None
    b = _dev(vk_bar, device) if vk_bar is not None else None
    out = torch.empty(nmol, nao, nao, dtype=torch.float64, device=device)
    w = _workspace(lib, device, nao, nmol)
    with torch.cuda.device(device):
        check(lib.qexxc_dot_eri_dm_vjp_batched(device.index, e.data_ptr(), a.data_ptr() if a is not None else None,
                                               b.data_ptr() if b is not None else None, nmol, nao, out.data_ptr(),
                                               w.data_ptr(), w.numel(), _stream(device)))
    return out


class _RowDotB(torch.autograd.Function):
    """Batched `J[b,i,j] = sum_kl eri[b,i,j,k,l] dm[b,k,l]` (see `_RowDot`)."""

    @staticmethod
    def forward(ctx, eri, dm):
        ctx.eri = eri
        B, nao = int(dm.shape[0]), int(dm.shape[-1])
        return _dot_eri_dm_batched_vjp(eri, nao, B, dm.contiguous(), None).transpose(-1, -2).contiguous()

    @staticmethod
    def backward(ctx, j_bar):
        vj, _ = dot_eri_dm_batched(ctx.eri, j_bar.transpose(-1, -2).contiguous(), True, False)
        return None, vj


def dot_eri_dm_rowdot_batched(eri, dm):
    return _RowDotB.apply(eri, dm)


def make_rdm1(mo_coeff, mo_occ):
    """hf_legacy.py:331-338: dm = (C_occ * occ) C_occ^T over orbitals with occ > 0."""
    c = torch.as_tensor(mo_coeff, dtype=torch.float64)
    o = torch.as_tensor(mo_occ, dtype=torch.float64, device=c.device)
    mask = o > 0
    mocc = c[:, mask]
    return (mocc * o[mask]) @ mocc.T


def energy_coulomb(dm, vj):
    """rks_legacy.py:122: 0.5 * einsum("ij,ji", dm, vj)."""
    return 0.5 * torch.sum(torch.as_tensor(dm) * torch.as_tensor(vj).transpose(-1, -2))


def jk_launch_count() -> int:
    return int(_lib.load().qexxc_jk_launch_count())
