"""Host mirror of ``qedft/train/td/numint_legacy.py`` for the accelerated branches.

``NumInt.nr_rks(mol, grids, xc_code, dms, relativity=0, hermi=0, max_memory=2000, verbose=None,
params=None) -> (nelec, excsum, vmat)`` (numint_legacy.py:122-348,588), ``eval_rho`` (:351-397) and
``eval_ao`` (pyscf ``numint.eval_ao`` as used at trainer_legacy_no_jit.py:273) with the same
argument meaning and error behaviour.  ``mol`` / ``grids`` are duck-typed (``mol._atm/_bas/_env``,
``mol.nao_nr()``, ``grids.coords``, ``grids.weights``), so pyscf objects work unchanged.

Two ways to evaluate the functional, as in the reference where ``ni.eval_xc`` is whatever
``define_xc_`` installed:
  * ``ni.eval_xc = qex_b200.xc.make_eval_xc(network, is_global_xc)``  -> the whole call is ONE
    fused device pass (``qexxc_nr_rks_fwd``), and ``nr_rks_vjp`` gives the reverse rule;
  * any other Python callable ``eval_xc(xc_code, rho, spin, relativity, deriv, verbose, params)``
    (the reference's tests use a toy ``0.01*rho**2``) -> rho is brought to the host for the callback
    and stages 2/4 still run on the device (``qexxc_eval_rho`` + ``qexxc_vxc_assemble``).
Differences from the reference, on purpose: no ``SWITCH_SIZE`` limit on nao (numint_legacy.py:451,465
raise ``NotImplementedError`` above it); AO values are evaluated on the device, not by pyscf.
"""
from __future__ import annotations

import hashlib
from functools import partial

import numpy as np

from . import _lib
from .engine import NetSpec, XCContext
from .xc import LDA_EXCHANGE_CODES, _native_apply, _theta, lda_exchange
from .xc import eval_xc as _native_eval_xc


class NPArrayWithTag(np.ndarray):
    """ndarray carrying attributes, like ``pyscf.lib.tag_array`` (how pyscf attaches ``mo_coeff`` /
    ``mo_occ`` to a density matrix; read by ``_gen_rho_evaluator``, numint_legacy.py:527-545)."""


def tag_array(a, **kwargs):
    t = np.asarray(a).view(NPArrayWithTag)
    t.__dict__.update(kwargs)
    return t


def _np(x):
    import torch

    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _xctype(ni, xc_code):
    """numint_legacy.py:133-140."""
    if isinstance(xc_code, str) and "NN" in xc_code:
        return xc_code
    if isinstance(xc_code, str) and xc_code.upper() in ("LDA", "GGA"):
        return xc_code.upper()
    if isinstance(xc_code, str) and xc_code.lower().replace(" ", "") in LDA_EXCHANGE_CODES:
        return "LDA"  # pyscf's _xc_type of Slater exchange
    t = getattr(ni, "_xc_type_override", None)
    if t is not None:
        return t
    raise NotImplementedError(
        f"xc_code {xc_code!r}: only 'NN', 'NN-AmplitudeEncoding', 'LDA' and 'GGA' (with a callable eval_xc) are on "
        "the accelerated path; libxc functionals stay with pyscf")


class NumInt:
    """Drop-in for ``numint_legacy.NumInt`` on the accelerated branches."""

    def __init__(self, device: int | None = None, cache_ao: bool = False):
        self.eval_xc = None  # installed by the caller / define_xc_
        self.device = device
        self.cache_ao = cache_ao
        self._ctxs = {}
        self._ao_key = {}

    # ---- context management ---------------------------------------------------------------
    def _native(self):
        fn = self.eval_xc
        if isinstance(fn, partial) and fn.func is _native_eval_xc:
            return _native_apply(fn.keywords["network"]), bool(fn.keywords.get("is_global_xc", True))
        return None, None

    def _ctx(self, nao, ngrids, ncomp, spec: NetSpec | None, nset: int = 1):
        """One cached context per (nao, ncomp, network, nset).  nset > 1 = several density matrices of one
        molecule: a batched context whose elements share the AO tensor (engine.XCContext(shared_ao=True))."""
        key = (nao, ncomp, None if spec is None else tuple(sorted(vars(spec).items())), nset)
        ctx = self._ctxs.get(key)
        if ctx is None or ctx.ngrids_max < ngrids:
            if ctx is not None:
                self._ao_key.pop(id(ctx), None)
                ctx.close()
            ctx = XCContext(nao=nao, ngrids_max=ngrids, ncomp=ncomp, nbatch=nset, net=spec, device=self.device,
                            shared_ao=nset > 1)
            self._ctxs[key] = ctx
            self._ao_key.pop(id(ctx), None)
        return ctx

    @staticmethod
    def _fingerprint(mol, coords, weights, deriv):
        """Content hash of everything stage 1 depends on (geometry + basis tables, grid, derivative order)."""
        h = hashlib.blake2b(digest_size=16)
        for a in (mol._atm, mol._bas, mol._env, coords, weights):
            a = np.ascontiguousarray(a)
            h.update(str((a.dtype.str, a.shape)).encode())
            h.update(a.tobytes())
        return h.hexdigest(), int(deriv)

    def _load(self, ctx, mol, grids, deriv):
        """Stage 1 on the device: grid upload + AO evaluation (re-done per call like the reference's
        block_loop, unless cache_ao and the same content fingerprint of mol tables + grid)."""
        coords, weights = np.asarray(grids.coords), np.asarray(grids.weights)
        fp = self._fingerprint(mol, coords, weights, deriv) if self.cache_ao else None
        if self.cache_ao and self._ao_key.get(id(ctx)) == fp:
            return
        ctx.set_grid(coords, weights)
        ctx.set_basis(mol._atm, mol._bas, mol._env)
        ctx.eval_ao(deriv)
        self._ao_key[id(ctx)] = fp

    # ---- B2 -------------------------------------------------------------------------------
    def eval_ao(self, mol, coords, deriv=0, **kwargs):
        """-> [G, N] (deriv=0) or [4, G, N] (deriv=1), float64 numpy."""
        if deriv not in (0, 1):
            raise NotImplementedError("eval_ao: deriv must be 0 or 1 on the accelerated path")
        coords = np.asarray(coords, dtype=np.float64)
        ncomp = 4 if deriv else 1
        ctx = self._ctx(mol.nao_nr(), coords.shape[0], ncomp, None)
        ctx.set_grid(coords, np.ones(coords.shape[0]))
        ctx.set_basis(mol._atm, mol._bas, mol._env)
        ctx.eval_ao(deriv)
        self._ao_key.pop(id(ctx), None)
        ao = ctx.get_ao(ncomp)[0].cpu().numpy()
        return ao[0] if deriv == 0 else ao

    def eval_rho(self, mol, ao, dm, non0tab=None, xctype="LDA", hermi=0, verbose=None):
        """numint_legacy.py:351-397: ao [G,N] (LDA) or [4,G,N] (GGA) -> rho [G] or [4,G]."""
        xctype = xctype.upper()
        if xctype in ("LDA", "HF"):
            ncomp = 1
        elif xctype in ("GGA", "NLC"):
            ncomp = 4
        else:
            raise NotImplementedError("eval_rho: meta-GGA is not on the accelerated path")
        ao = _np(ao)
        G, N = ao.shape[-2], ao.shape[-1]
        ctx = self._ctx(N, G, ncomp, None)
        ctx.set_grid(None, np.ones(G))
        ctx.set_ao(ao, ncomp)
        self._ao_key.pop(id(ctx), None)
        rho = ctx.eval_rho(_np(dm), ncomp, hermi)[0].cpu().numpy()
        return rho[0] if ncomp == 1 else rho

    def eval_mat(self, mol, ao, weight, rho, vxc, non0tab=None, xctype="LDA", spin=0, verbose=None):
        """numint_legacy.py:23-120 (closed-shell LDA / GGA): the V_xc matrix from caller-supplied
        potentials, `mat + mat.T` with the same 0.5*w*vrho and 2*w*vsigma*grad(rho) scale factors."""
        xctype = xctype.upper()
        if spin != 0:
            raise NotImplementedError("eval_mat: spin-polarised assembly is not on the accelerated path")
        if xctype not in ("LDA", "HF", "GGA"):
            raise NotImplementedError("eval_mat: meta-GGA is not on the accelerated path")
        ao = _np(ao)
        gga = xctype == "GGA"
        G, N = ao.shape[-2], ao.shape[-1]
        if gga:
            vrho, vsigma = vxc[:2]
            rho = _np(rho)
            assert vsigma is not None and rho.ndim == 2
        else:
            vrho = vxc[0] if not getattr(vxc, "ndim", None) == 2 else vxc
            vsigma = None
            rho = np.zeros(G) if rho is None else _np(rho).reshape(-1)[:G]
        ncomp = 4 if gga else 1
        ctx = self._ctx(N, G, ncomp, None)
        ctx.set_grid(None, _np(weight))
        ctx.set_ao(ao[:4] if gga else ao, ncomp)
        self._ao_key.pop(id(ctx), None)
        out = ctx.vxc_assemble(rho[:4] if gga else rho, np.zeros(G), _np(vrho).reshape(-1),
                               None if vsigma is None else _np(vsigma).reshape(-1), "GGA" if gga else "NN")
        return out[0, : N * N].reshape(N, N).cpu().numpy()

    # ---- B1 -------------------------------------------------------------------------------
    def nr_rks(self, mol, grids, xc_code, dms, relativity=0, hermi=0, max_memory=2000, verbose=None, params=None,
               return_resid=False):
        """numint_legacy.py:122-348.  `dms` [N,N] or [nset,N,N]; the results are unwrapped when nset == 1
        (:344-348).  The nset density matrices of one call go through ONE batched launch per stage (they share
        the AO tensor), not a host loop.  With `return_resid`, a fourth value is appended: the residual buffer of
        the whole call (all nset density matrices), to be handed to `nr_rks_vjp`."""
        xctype = _xctype(self, xc_code)
        dms_in = dms
        dms = _np(dms)
        dm_arr = dms[None] if dms.ndim == 2 else dms
        nset = dm_arr.shape[0]
        fn, is_global = self._native()
        gga = xctype == "GGA"
        ncomp = 4 if gga else 1
        N, G = mol.nao_nr(), int(np.asarray(grids.weights).shape[0])
        kind = "NN" if xctype == "LDA" else xctype
        # numint_legacy.py:527-545: a dm tagged with mo_coeff / mo_occ takes the MO form of rho (eval_rho2)
        mo_coeff, mo_occ = getattr(dms_in, "mo_coeff", None), getattr(dms_in, "mo_occ", None)
        use_mo = mo_coeff is not None and mo_occ is not None and nset == 1 and not gga
        if use_mo:
            mo_coeff, mo_occ = _np(mo_coeff), _np(mo_occ)
            keep = np.abs(mo_occ) > 1e-12  # OCCDROP: only occupied orbitals enter (pos/neg handled by sign)
            mo_coeff, mo_occ = np.ascontiguousarray(mo_coeff[:, keep]), np.ascontiguousarray(mo_occ[keep])
            use_mo = mo_occ.size > 0
        if fn is not None and xctype == "NN-AmplitudeEncoding" and not is_global:
            raise ValueError("xctype 'NN-AmplitudeEncoding' needs eval_xc built with is_global_xc=True")
        # a local network under is_global_xc=True (the reference's default flag) is the sum of its per-point
        # outputs (trainer_legacy_no_jit.py:46-53): no fused kernel for that pairing, so it runs as
        # eval_rho -> eval_xc -> vxc_assemble on the device buffers below
        fused = fn is not None and not (is_global and fn.qex_spec.kind != _lib.NET_GLOBAL_MLP)
        resid = None
        if fused:
            ctx = self._ctx(N, G, ncomp, fn.qex_spec, nset)
            self._load(ctx, mol, grids, 1 if gga else 0)
            theta = _theta(fn, params)
            if use_mo:
                out, resid = ctx.nr_rks_fwd_mo(mo_coeff, mo_occ, theta, kind, want_resid=return_resid)
            else:
                out, resid = ctx.nr_rks_fwd(dm_arr, theta, kind, hermi, want_resid=return_resid)
        else:
            builtin_lda = not callable(self.eval_xc) and xctype == "LDA"
            if not callable(self.eval_xc) and not builtin_lda:
                raise ValueError("NumInt.eval_xc is not set: install a functional first (define_xc_)")
            ctx = self._ctx(N, G, ncomp, None, nset)
            self._load(ctx, mol, grids, 1 if gga else 0)
            rho = ctx.eval_rho_mo(mo_coeff, mo_occ) if use_mo else ctx.eval_rho(dm_arr, ncomp, hermi)
            if builtin_lda:
                # the reference's libxc route for xc_code "lda" (Slater exchange), all on the device
                exc_d, vrho_d = lda_exchange(rho[:, 0, :])
                out = ctx.vxc_assemble(rho, exc_d, vrho_d, None, "NN")
            else:
                r = rho.cpu().numpy()
                excs, vrhos, vgams = [], [], []
                for i in range(nset):  # host callback per density matrix, as the reference's `for idm in range(nset)`
                    exc, vxc = self.eval_xc(xc_code, r[i, 0] if ncomp == 1 else r[i], spin=0, relativity=relativity,
                                            deriv=1, verbose=verbose, params=params)[:2]
                    excs.append(np.atleast_1d(_np(exc)))
                    vrhos.append(_np(vxc[0]))
                    if gga:
                        vgams.append(_np(vxc[1]))
                akind = "NN-AmplitudeEncoding" if xctype == "NN-AmplitudeEncoding" else ("GGA" if gga else "NN")
                out = ctx.vxc_assemble(rho, np.stack(excs), np.stack(vrhos), np.stack(vgams) if gga else None, akind)
        o = out.cpu().numpy()  # the one host synchronisation of the call
        vmat = [o[i, : N * N].reshape(N, N).copy() for i in range(nset)]
        excsum = [float(o[i, N * N]) for i in range(nset)]
        nelec = [float(o[i, N * N + 1]) for i in range(nset)]
        if nset == 1:  # numint_legacy.py:344-348
            res = (nelec[0], excsum[0], vmat[0])
        else:
            res = (nelec, excsum, vmat)
        return res + (resid,) if return_resid else res

    def nr_rks_vjp(self, mol, grids, xc_code, resid, e_bar, v_bar, hermi=0, params=None):
        """Reverse rule of ``nr_rks`` (native functional only): cotangents of (excsum, vmat) ->
        (dm_bar [N,N], theta_bar flat); for nset > 1: e_bar [nset], v_bar [nset,N,N] -> dm_bar [nset,N,N] and
        theta_bar summed over the set.  ``resid`` comes from ``nr_rks(..., return_resid=True)``."""
        fn, _ = self._native()
        if fn is None:
            raise NotImplementedError("nr_rks_vjp needs a native functional (xc.make_eval_xc)")
        xctype = _xctype(self, xc_code)
        gga = xctype == "GGA"
        N, G = mol.nao_nr(), int(np.asarray(grids.weights).shape[0])
        v_bar = _np(v_bar)
        nset = 1 if v_bar.ndim == 2 else v_bar.shape[0]
        ctx = self._ctx(N, G, 4 if gga else 1, fn.qex_spec, nset)
        self._load(ctx, mol, grids, 1 if gga else 0)
        theta = _theta(fn, params)
        bar = ctx.nr_rks_vjp(theta, resid, np.atleast_1d(np.asarray(e_bar, dtype=np.float64)), v_bar,
                             "NN" if xctype == "LDA" else xctype, hermi).cpu().numpy()
        dm_bar = bar[: nset * N * N].reshape(nset, N, N)
        return (dm_bar[0] if nset == 1 else dm_bar), bar[nset * N * N :]


_default = None


def _ni():
    global _default
    if _default is None:
        _default = NumInt()
    return _default


def eval_ao(mol, coords, deriv=0, **kwargs):
    """``pyscf.dft.numint.eval_ao`` as the trainer calls it (trainer_legacy_no_jit.py:273)."""
    return _ni().eval_ao(mol, coords, deriv, **kwargs)


def eval_rho(mol, ao, dm, non0tab=None, xctype="LDA", hermi=0, verbose=None):
    """numint_legacy.py:351."""
    return _ni().eval_rho(mol, ao, dm, non0tab, xctype, hermi, verbose)


def eval_mat(mol, ao, weight, rho, vxc, non0tab=None, xctype="LDA", spin=0, verbose=None):
    return _ni().eval_mat(mol, ao, weight, rho, vxc, non0tab, xctype, spin, verbose)


def nr_rks(ni, mol, grids, xc_code, dms, relativity=0, hermi=0, max_memory=2000, verbose=None, params=None):
    """numint_legacy.py:122 (free-function form; ``NumInt.nr_rks = nr_rks`` at :588)."""
    return ni.nr_rks(mol, grids, xc_code, dms, relativity, hermi, max_memory, verbose, params)
