"""Thin object wrapper over the C ABI: one ``XCContext`` per (device, nao, ngrids_max, network).

torch is plumbing here (device buffers, the current CUDA stream); every number is produced by
the kernels in libqexxc.so.  All methods take / return float64 CUDA tensors; numpy inputs are
copied to the device first.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import (ACTIVATIONS, NET_GLOBAL_MLP, NET_LOCAL_MLP, NET_LOCAL_QNN, NET_NONE, PREC_F32, PREC_F64, XC_GGA,
                   XC_NN, XC_NN_GLOBAL, NetDesc, check)

XCTYPES = {"NN": XC_NN, "LDA": XC_NN, "NN-AmplitudeEncoding": XC_NN_GLOBAL, "GGA": XC_GGA}


@dataclass
class NetSpec:
    """Host description of the XC network (mirrors qexxc_net_desc)."""

    kind: int = NET_NONE
    n_features: int = 1
    n_hidden: int = 3
    width: int = 64
    activation: str = "tanh"
    out_transform: int = 0
    precision: str = "f64"
    in_scale: float = 0.5
    out_scale: float = 1e-2

    def desc(self) -> NetDesc:
        if self.activation not in ACTIVATIONS:
            raise ValueError(f"Unknown activation '{self.activation}'. Valid options: {list(ACTIVATIONS)}")
        return NetDesc(self.kind, self.n_features, self.n_hidden, self.width, ACTIVATIONS[self.activation],
                       self.out_transform, PREC_F32 if self.precision == "f32" else PREC_F64, 0, self.in_scale,
                       self.out_scale)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _xct(xctype) -> int:
    if isinstance(xctype, str):
        if xctype not in XCTYPES:
            raise NotImplementedError(f"xctype {xctype!r} is not on the accelerated path")
        return XCTYPES[xctype]
    return int(xctype)


class XCContext:
    def __init__(self, nao: int, ngrids_max: int, ncomp: int = 1, nbatch: int = 1, net: NetSpec | None = None,
                 device: int | None = None, shared_ao: bool = False):
        """shared_ao: the batch is `nset` density matrices of ONE molecule (numint_legacy.py:141-156): a single
        AO tensor / grid / geometry serves all batch elements (QEXXC_FLAG_SHARED_AO)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.QexxcError(_lib.ERR_NODEVICE, "no CUDA device: qex_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        self.nao, self.ngrids_max, self.ncomp, self.nbatch = int(nao), int(ngrids_max), int(ncomp), int(nbatch)
        self.net = net or NetSpec()
        self._desc = self.net.desc()
        self._h = C.c_void_p()
        self.shared_ao = bool(shared_ao)
        self._nb_geom = 1 if self.shared_ao else self.nbatch
        check(self.lib.qexxc_create_ex(C.byref(self._h), self.device, self.nbatch, self.ncomp, self.ngrids_max,
                                       self.nao, C.byref(self._desc), 1 if self.shared_ao else 0))
        self.ngrids = 0
        self.n_params = int(self.lib.qexxc_n_params(C.byref(self._desc), self.ngrids_max))

    # ---- plumbing --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.qexxc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dev(self, x, shape=None) -> torch.Tensor:
        """float64 contiguous tensor on this context's device."""
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.tdev, dtype=torch.float64)
        else:
            t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(self.tdev)
        t = t.contiguous()
        if shape is not None:
            if t.numel() != int(np.prod(shape)):
                raise ValueError(f"expected {int(np.prod(shape))} elements for shape {tuple(shape)}, got {tuple(t.shape)}")
            t = t.reshape(shape)
        return t

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float64, device=self.tdev)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.qexxc_workspace_bytes(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.qexxc_launch_count(self._h))

    @property
    def resid_doubles(self) -> int:
        return int(self.lib.qexxc_resid_doubles(self._h))

    # ---- stage 1 ---------------------------------------------------------------------------
    def set_grid(self, coords, weights):
        B = self._nb_geom
        w = self.dev(weights)
        G = w.numel() // B
        w = w.reshape(B, G)
        c = self.dev(coords, (B, G, 3)) if coords is not None else None
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_set_grid(self._h, self._p(c), self._p(w), G, _stream()))
        self.ngrids = G
        self._keep = (c, w)
        return self

    def set_basis(self, atm, bas, env):
        atm = np.ascontiguousarray(atm, dtype=np.int32).reshape(-1, 6)
        bas = np.ascontiguousarray(bas, dtype=np.int32).reshape(-1, 8)
        env = np.ascontiguousarray(env, dtype=np.float64)
        env = np.broadcast_to(env, (self._nb_geom, env.shape[-1])) if env.ndim == 1 else env
        env = np.ascontiguousarray(env)
        if env.shape[0] != self._nb_geom:
            raise ValueError("env must be [nenv] or [nbatch, nenv]")
        check(self.lib.qexxc_set_basis(self._h, atm.ctypes.data_as(C.c_void_p), atm.shape[0],
                                       bas.ctypes.data_as(C.c_void_p), bas.shape[0],
                                       env.ctypes.data_as(C.c_void_p), env.shape[1]))
        return self

    def eval_ao(self, deriv: int = 0):
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_eval_ao(self._h, int(deriv), _stream()))
        return self

    def set_ao(self, ao, ncomp: int | None = None):
        B, G, N = self._nb_geom, self.ngrids, self.nao
        a = self.dev(ao)
        if ncomp is None:
            ncomp = a.numel() // (B * G * N)
        a = a.reshape(B, ncomp, G, N)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_set_ao(self._h, self._p(a), ncomp, G, _stream()))
        return self

    def get_ao(self, ncomp: int = 1) -> torch.Tensor:
        out = self.empty(self._nb_geom, ncomp, self.ngrids, self.nao)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_get_ao(self._h, self._p(out), ncomp, _stream()))
        return out

    # ---- stage 2 ---------------------------------------------------------------------------
    def eval_rho(self, dm, ncomp: int = 1, hermi: int = 0) -> torch.Tensor:
        d = self.dev(dm, (self.nbatch, self.nao, self.nao))
        out = self.empty(self.nbatch, ncomp, self.ngrids)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_eval_rho(self._h, self._p(d), ncomp, int(hermi), self._p(out), _stream()))
        return out

    def eval_rho_vjp(self, rho_bar, ncomp: int = 1, hermi: int = 0) -> torch.Tensor:
        rb = self.dev(rho_bar, (self.nbatch, ncomp, self.ngrids))
        out = self.empty(self.nbatch, self.nao, self.nao)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_eval_rho_vjp(self._h, self._p(rb), ncomp, int(hermi), self._p(out), _stream()))
        return out

    # ---- stage 3 ---------------------------------------------------------------------------
    def xc_fwd(self, rho, theta, xctype="NN"):
        xt = _xct(xctype)
        B, G = self.nbatch, self.ngrids
        nc = 4 if xt == XC_GGA else 1
        r = self.dev(rho, (B, nc, G))
        th = self.dev(theta)
        exc = self.empty(B) if xt == XC_NN_GLOBAL else self.empty(B, G)
        vrho = self.empty(B, G)
        vgamma = self.empty(B, G) if xt == XC_GGA else None
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_xc_fwd(self._h, xt, self._p(r), self._p(th), th.numel(), self._p(exc), self._p(vrho),
                                        self._p(vgamma), _stream()))
        return exc, vrho, vgamma

    def xc_vjp(self, rho, theta, exc_bar, vrho_bar, vgamma_bar=None, xctype="NN"):
        xt = _xct(xctype)
        B, G = self.nbatch, self.ngrids
        nc = 4 if xt == XC_GGA else 1
        r = self.dev(rho, (B, nc, G))
        th = self.dev(theta)
        eb = self.dev(exc_bar, (B,) if xt == XC_NN_GLOBAL else (B, G))
        vb = self.dev(vrho_bar, (B, G))
        gb = self.dev(vgamma_bar, (B, G)) if xt == XC_GGA else None
        rbar = self.empty(B, nc, G)
        tbar = self.empty(th.numel())
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_xc_vjp(self._h, xt, self._p(r), self._p(th), th.numel(), self._p(eb), self._p(vb), self._p(gb),
                                        self._p(rbar), self._p(tbar), _stream()))
        return rbar, tbar

    def apply_fn(self, x, theta) -> torch.Tensor:
        xx = self.dev(x)
        th = self.dev(theta)
        if self.net.kind == NET_GLOBAL_MLP:
            npts, y = xx.numel(), self.empty(1)
        else:
            F = self.net.n_features if self.net.kind == NET_LOCAL_MLP else 1
            npts = xx.numel() // F
            y = self.empty(npts)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_apply_fn_fwd(self._h, self._p(xx), npts, self._p(th), th.numel(), self._p(y), _stream()))
        return y

    def apply_fn_vjp(self, x, theta, y_bar):
        xx = self.dev(x)
        th = self.dev(theta)
        yb = self.dev(y_bar)
        if self.net.kind == NET_GLOBAL_MLP:
            npts = xx.numel()
        else:
            F = self.net.n_features if self.net.kind == NET_LOCAL_MLP else 1
            npts = xx.numel() // F
        xb = torch.empty_like(xx)
        tb = self.empty(th.numel())
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_apply_fn_vjp(self._h, self._p(xx), npts, self._p(th), th.numel(), self._p(yb), self._p(xb),
                                              self._p(tb), _stream()))
        return xb, tb

    # ---- stage 4 ---------------------------------------------------------------------------
    def vxc_assemble(self, rho, exc, vrho, vgamma=None, xctype="NN") -> torch.Tensor:
        xt = _xct(xctype)
        B, G, N = self.nbatch, self.ngrids, self.nao
        nc = 4 if xt == XC_GGA else 1
        r = self.dev(rho, (B, nc, G))
        e = self.dev(exc, (B,) if xt == XC_NN_GLOBAL else (B, G))
        v = self.dev(vrho, (B, G))
        g = self.dev(vgamma, (B, G)) if xt == XC_GGA else None
        out = self.empty(B, N * N + 2)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_vxc_assemble(self._h, xt, self._p(r), self._p(e), self._p(v), self._p(g),
                                              self._p(out), _stream()))
        return out

    def vxc_assemble_vjp(self, rho, exc, vrho, vgamma, e_bar, v_bar, xctype="NN"):
        xt = _xct(xctype)
        B, G, N = self.nbatch, self.ngrids, self.nao
        nc = 4 if xt == XC_GGA else 1
        r = self.dev(rho, (B, nc, G))
        e = self.dev(exc, (B,) if xt == XC_NN_GLOBAL else (B, G))
        v = self.dev(vrho, (B, G))
        g = self.dev(vgamma, (B, G)) if xt == XC_GGA else None
        eb = self.dev(e_bar, (B,))
        vb = self.dev(v_bar, (B, N, N))
        rbar = self.empty(B, nc, G)
        excb = self.empty(B) if xt == XC_NN_GLOBAL else self.empty(B, G)
        vrb = self.empty(B, G)
        vgb = self.empty(B, G) if xt == XC_GGA else None
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_vxc_assemble_vjp(self._h, xt, self._p(r), self._p(e), self._p(v), self._p(g),
                                                  self._p(eb), self._p(vb), self._p(rbar), self._p(excb),
                                                  self._p(vrb), self._p(vgb), _stream()))
        return rbar, excb, vrb, vgb

    # ---- fused hot path --------------------------------------------------------------------
    def nr_rks_fwd(self, dm, theta, xctype="NN", hermi: int = 0, want_resid: bool = True, out=None, resid=None):
        """-> (out [B, N*N+2] = vmat | excsum | nelec, resid or None)."""
        xt = _xct(xctype)
        B, N = self.nbatch, self.nao
        d = self.dev(dm, (B, N, N))
        th = self.dev(theta)
        if out is None:
            out = self.empty(B, N * N + 2)
        if want_resid and resid is None:
            resid = self.empty(self.resid_doubles)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_nr_rks_fwd(self._h, xt, int(hermi), self._p(d), self._p(th), th.numel(), self._p(out),
                                            self._p(resid if want_resid else None), _stream()))
        return out, (resid if want_resid else None)

    def _mo(self, mo_coeff, mo_occ):
        B, N = self.nbatch, self.nao
        occ = self.dev(mo_occ).reshape(B, -1)
        nmo = occ.shape[1]
        return self.dev(mo_coeff, (B, N, nmo)), occ, nmo

    def eval_rho_mo(self, mo_coeff, mo_occ) -> torch.Tensor:
        """rho = sum_k occ_k (ao C_k)^2 (pyscf eval_rho2, numint_legacy.py:527-545) -> [B, 1, G]."""
        C_, occ, nmo = self._mo(mo_coeff, mo_occ)
        out = self.empty(self.nbatch, 1, self.ngrids)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_eval_rho_mo(self._h, self._p(C_), self._p(occ), nmo, self._p(out), _stream()))
        return out

    def nr_rks_fwd_mo(self, mo_coeff, mo_occ, theta, xctype="NN", want_resid: bool = True, out=None, resid=None):
        """nr_rks forward with the MO form of stage 2 (dm = C occ C^T is never formed)."""
        xt = _xct(xctype)
        B, N = self.nbatch, self.nao
        C_, occ, nmo = self._mo(mo_coeff, mo_occ)
        th = self.dev(theta)
        if out is None:
            out = self.empty(B, N * N + 2)
        if want_resid and resid is None:
            resid = self.empty(self.resid_doubles)
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_nr_rks_fwd_mo(self._h, xt, self._p(C_), self._p(occ), nmo, self._p(th), th.numel(), self._p(out),
                                               self._p(resid if want_resid else None), _stream()))
        return out, (resid if want_resid else None)

    def nr_rks_vjp(self, theta, resid, e_bar, v_bar, xctype="NN", hermi: int = 0, out=None):
        """-> bar [B*N*N + n_theta] = dm_bar | theta_bar."""
        xt = _xct(xctype)
        B, N = self.nbatch, self.nao
        th = self.dev(theta)
        eb = self.dev(e_bar, (B,))
        vb = self.dev(v_bar, (B, N, N))
        if out is None:
            out = self.empty(B * N * N + th.numel())
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_nr_rks_vjp(self._h, xt, int(hermi), self._p(th), th.numel(), self._p(resid), self._p(eb),
                                            self._p(vb), self._p(out), _stream()))
        return out

    def debug_run_contraction(self, which: int):
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_debug_run_contraction(self._h, int(which), _stream()))

    def contraction_flops(self, which: int, symmetric: bool) -> float:
        """DMMA FLOPs one launch of rowquad (0) / wsyrk (1) executes for the current shape."""
        v = C.c_double(0)
        check(self.lib.qexxc_contraction_flops(self._h, int(which), 1 if symmetric else 0, C.byref(v)))
        return v.value

    @property
    def contraction_mode(self) -> str:
        """"int8" (exact digit split on tcgen05, contract_i8.cu) or "dmma" (FP64 tensor pipe) for this context."""
        return "int8" if self.lib.qexxc_contraction_mode(self._h) else "dmma"

    def prepare_contractions(self):
        """Build what the contractions derive from the AO tensor alone (INT8 mode: the digit planes) now, on the current
        stream, instead of inside the first contraction -- lets it overlap an upload of the density matrix."""
        with torch.cuda.device(self.device):
            check(self.lib.qexxc_prepare_contractions(self._h, _stream()))
        return self

    def contraction_i8_ops(self, which: int, symmetric: bool) -> float:
        """INT8 operations (2 per MAC) one launch of rowquad (0) / wsyrk (1) executes in "int8" mode."""
        v = C.c_double(0)
        check(self.lib.qexxc_contraction_i8_ops(self._h, int(which), 1 if symmetric else 0, C.byref(v)))
        return v.value

    PROF_CLASSES = {"rowquad": 0, "wsyrk": 1, "xc_fwd": 2, "xc_vjp": 3, "eval_ao": 4, "stage4": 5, "slice": 6}

    def profile_enable(self, on: bool = True):
        check(self.lib.qexxc_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> dict:
        """{class: (total_ms, launches)} since the last read (synchronises on the recorded events)."""
        out = {}
        for name, k in self.PROF_CLASSES.items():
            ms, n = C.c_double(0), C.c_long(0)
            check(self.lib.qexxc_profile_read(self._h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out
