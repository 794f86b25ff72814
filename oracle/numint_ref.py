"""Oracle: XC grid integration (stages 2 and 4) forward and reverse.  TEST INFRASTRUCTURE ONLY.

NumPy float64 restatement of ``qedft/train/td/numint_legacy.py`` (pasqal-io/qex):

* ``eval_rho``           <- numint_legacy.py:351-397 (+ ``_rks_gga_assemble_rho`` :401-410,
                            ``_dot_ao_dm_incore`` :469-471, ``_contract_rho`` :475-481)
* ``scale_ao``           <- ``_scale_ao`` :432-442
* ``rks_gga_wv0``        <- ``_rks_gga_wv0`` :485-491
* ``nr_rks``             <- ``nr_rks`` :122-348, branches "NN" :290-310, "NN-AmplitudeEncoding"
                            :311-334, "GGA" :175-198; dm symmetrisation from
                            ``NumInt._gen_rho_evaluator`` :548-558; final ``vmat + vmat.T`` :336-337
* ``nr_rks_vjp``         <- the reverse-mode rule JAX derives for the above under
                            ``jax.value_and_grad`` (``trainer_legacy_no_jit.py:284``); hand-derived
                            here and checked by finite differences.
* ``get_veff_xc_einsum`` <- ``scf_functions_masked.py:143-159`` (second statement of the same math).

The block loop over the grid (pyscf ``block_loop``) is restated with a fixed block size: the
reference accumulates ``nelec``, ``excsum`` and ``vmat`` block by block (:305-309).

Pinned by: the closed-form toy functional of the reference's own tests
(``tests/test_numint.py:96-103``), the einsum cross-statement, finite differences.
"""
from __future__ import annotations

import numpy as np

BLKSIZE = 128  # pyscf.dft.numint.BLKSIZE (2.9: 56 or 128 depending on build; only affects rounding)


def _blocks(G, blk):
    for s in range(0, G, blk):
        yield slice(s, min(G, s + blk))


def _default_block(G, nao, ncomp, max_memory=2000):
    # pyscf NumInt.block_loop: blksize = int(max_memory*1e6/((comp+1)*nao*8*BLKSIZE)) * BLKSIZE,
    # clipped to [BLKSIZE, ngrids]
    blk = int(max_memory * 1e6 / ((ncomp + 1) * max(nao, 1) * 8 * BLKSIZE))
    blk = max(1, min(blk, (G + BLKSIZE - 1) // BLKSIZE)) * BLKSIZE
    return max(BLKSIZE, blk)


def contract_rho(bra, ket, factor=1.0):
    """numint_legacy.py:475-481 (real inputs)."""
    return np.einsum("pi,pi->p", bra, ket) * factor


def eval_rho(ao, dm, xctype="LDA", hermi=0):
    """numint_legacy.py:351-397.  ao [G,N] (LDA) or [4,G,N] (GGA); returns [G] or [4,G]."""
    xctype = xctype.upper()
    dm = np.asarray(dm, dtype=np.float64)
    if not hermi:
        dm = (dm + dm.T) * 0.5  # :362-365
    if xctype in ("LDA", "HF"):
        c0 = ao @ dm  # _dot_ao_dm_incore
        return contract_rho(ao, c0)
    if xctype in ("GGA", "NLC"):
        c0 = ao[0] @ dm  # _rks_gga_assemble_rho :401-410
        rho = [contract_rho(c0, ao[0], 1.0)]
        for i in range(1, 4):
            rho.append(contract_rho(c0, ao[i], 2.0))
        return np.asarray(rho)
    raise NotImplementedError("oracle eval_rho: xctype " + xctype)


def eval_rho2(ao, mo_coeff, mo_occ, xctype="LDA"):
    """pyscf ``numint.eval_rho2`` (LDA part), reached through ``NumInt._gen_rho_evaluator`` when the
    density matrix carries ``mo_coeff`` / ``mo_occ`` (numint_legacy.py:527-545):
    rho = sum_k occ_k (ao C_k)^2 for occ_k > OCCDROP, minus the same for occ_k < -OCCDROP."""
    if xctype.upper() not in ("LDA", "HF"):
        raise NotImplementedError("oracle eval_rho2: LDA only")
    OCCDROP = 1e-12
    mo_coeff, mo_occ = np.asarray(mo_coeff, dtype=np.float64), np.asarray(mo_occ, dtype=np.float64)
    rho = np.zeros(ao.shape[0])
    pos = mo_occ > OCCDROP
    if pos.sum() > 0:
        c0 = ao @ (mo_coeff[:, pos] * np.sqrt(mo_occ[pos]))
        rho += contract_rho(c0, c0)
    neg = mo_occ < -OCCDROP
    if neg.sum() > 0:
        c0 = ao @ (mo_coeff[:, neg] * np.sqrt(-mo_occ[neg]))
        rho -= contract_rho(c0, c0)
    return rho


def scale_ao(ao, wv):
    """numint_legacy.py:432-442: aow[p,i] = sum_n ao[n,p,i] wv[n,p]."""
    if wv.ndim == 2:
        return np.einsum("npi,np->pi", ao[: wv.shape[0]], wv)
    return ao * wv[:, None]


def rks_gga_wv0(rho, vxc, weight):
    """numint_legacy.py:485-491."""
    vrho, vgamma = vxc[:2]
    wv_rho = weight * vrho * 0.5
    wv_sigma = (weight * vgamma * 2) * rho[1:4]
    return np.concatenate((wv_rho.reshape(1, -1), wv_sigma))


def nr_rks(ao, weights, dm, eval_xc, xctype="NN", hermi=0, params=None, blksize=None):
    """numint_legacy.py:122-348 for one density matrix (nset == 1).

    ``ao``: [G,N] for "NN"/"NN-AmplitudeEncoding"/"LDA", [4,G,N] for "GGA".
    ``eval_xc(xc_code, rho, spin=0, relativity=0, deriv=1, verbose=None, params=params)``
    returns ``(exc, (vrho, vgamma, vlapl, vtau), fxc, kxc)`` as at numint_legacy.py:295-303.
    Returns ``(nelec, excsum, vmat)``.
    """
    dm = np.asarray(dm, dtype=np.float64)
    if not hermi:
        dm = (dm + dm.T) * 0.5  # _gen_rho_evaluator :551-553
    gga = xctype == "GGA"
    G = ao.shape[1] if gga else ao.shape[0]
    N = ao.shape[-1]
    if blksize is None:
        blksize = G if xctype == "NN-AmplitudeEncoding" else _default_block(G, N, 4 if gga else 1)
    nelec = 0.0
    excsum = 0.0
    vmat = np.zeros((N, N))
    for sl in _blocks(G, blksize):
        w = weights[sl]
        if gga:
            a = ao[:, sl]
            rho = eval_rho(a, dm, "GGA", hermi=1)
            exc, vxc = eval_xc(xctype, rho, spin=0, relativity=0, deriv=1, verbose=None, params=params)[:2]
            den = rho[0] * w
            nelec += den.sum()
            excsum += np.dot(den, exc)
            wv = rks_gga_wv0(rho, vxc, w)
            aow = scale_ao(a, wv)
            vmat += a[0].T @ aow
        else:
            a = ao[sl]
            rho = eval_rho(a, dm, "LDA", hermi=1)
            exc, vxc = eval_xc(xctype, rho, spin=0, relativity=0, deriv=1, verbose=None, params=params)[:2]
            vrho = vxc[0]
            den = rho * w
            nelec += den.sum()
            if xctype == "NN-AmplitudeEncoding":
                excsum += exc  # :331 -- the network's scalar, no grid weights
            else:
                excsum += np.dot(den, exc)  # :306
            aow = scale_ao(a, 0.5 * w * vrho)  # :308
            vmat += a.T @ aow  # :309
    vmat = vmat + vmat.T  # :336-337
    return nelec, excsum, vmat


def get_veff_xc_einsum(ao, weights, dm, eval_xc, params=None):
    """scf_functions_masked.py:143-159 (XC part): einsum statement, equal to ``nr_rks`` "NN"
    when dm is symmetric."""
    rho = np.einsum("gi,ij,gj->g", ao, dm, ao)
    exc, (vrho, _, _, _), _, _ = eval_xc("", rho, params=params)
    vxc = np.einsum("gi,g,gj->ij", ao, weights * vrho, ao)
    return vxc, np.sum(exc * rho * weights)


# ----------------------------------------------------------------------------------------
# reverse mode
# ----------------------------------------------------------------------------------------
def nr_rks_vjp(ao, weights, dm, xc_fwd, xc_vjp, e_bar, v_bar, xctype="NN", hermi=0, params=None):
    """Reverse-mode rule of ``nr_rks`` w.r.t. (dm, params); ``nelec`` is stop-gradient
    (numint_legacy.py:305).  Not blocked (the result is block-order independent up to rounding).

    ``xc_fwd(rho, params) -> (exc, vrho[, vgamma])``
    ``xc_vjp(rho, params, exc_bar, vrho_bar[, vgamma_bar]) -> (rho_bar_like_inputs, params_bar)``
      * local ("NN"): rho [G]; exc_bar [G]; returns rho_bar [G]
      * global ("NN-AmplitudeEncoding"): exc scalar; exc_bar scalar
      * "GGA" (extension): xc works on features (rho0, sigma); returns
        (rho0_bar [G], sigma_bar [G])
    Returns ``(dm_bar [N,N], params_bar)``.
    """
    dm = np.asarray(dm, dtype=np.float64)
    dms = (dm + dm.T) * 0.5 if not hermi else dm
    M = v_bar + v_bar.T  # adjoint of vmat + vmat.T
    w = weights
    if xctype == "GGA":
        a0 = ao[0]
        rho = eval_rho(ao, dms, "GGA", hermi=1)
        sigma = (rho[1:4] ** 2).sum(0)
        exc, vrho, vgamma = xc_fwd(np.stack([rho[0], sigma]), params)
        u = a0 @ M  # adjoint of vmat_half = a0.T @ aow -> aow_bar = a0 @ M ... (M symmetric)
        wv_bar = np.stack([np.einsum("gj,gj->g", ao[c], u) for c in range(4)])
        # aow = sum_c ao[c] * wv[c]; vmat_half = a0.T @ aow; also a0 appears on the left:
        # its adjoint does not flow anywhere (AO are not differentiated).
        vrho_bar = 0.5 * w * wv_bar[0]
        vgamma_bar = 2.0 * w * (rho[1:4] * wv_bar[1:4]).sum(0)
        rho_bar = np.zeros_like(rho)
        rho_bar[1:4] += 2.0 * w * vgamma * wv_bar[1:4]
        exc_bar = e_bar * w * rho[0]
        rho_bar[0] += e_bar * w * exc
        (r0_bar, sig_bar), p_bar = xc_vjp(np.stack([rho[0], sigma]), params, exc_bar, vrho_bar, vgamma_bar)
        rho_bar[0] += r0_bar
        rho_bar[1:4] += sig_bar * 2.0 * rho[1:4]
        # rho_c = f_c * rowdot(a0 @ dms, ao[c]); f = (1,2,2,2)
        f = np.array([1.0, 2.0, 2.0, 2.0])
        t = np.einsum("c,cg,cgj->gj", f, rho_bar, ao)
        D = a0.T @ t
    else:
        rho = eval_rho(ao, dms, "LDA", hermi=1)
        u = ao @ M
        wv_bar = np.einsum("gj,gj->g", ao, u)
        vrho_bar = 0.5 * w * wv_bar
        if xctype == "NN-AmplitudeEncoding":
            exc, vrho = xc_fwd(rho, params)
            rho_bar, p_bar = xc_vjp(rho, params, e_bar, vrho_bar)
        else:
            exc, vrho = xc_fwd(rho, params)
            exc_bar = e_bar * w * rho
            rho_bar, p_bar = xc_vjp(rho, params, exc_bar, vrho_bar)
            rho_bar = rho_bar + e_bar * w * exc
        # rho = rowdot(ao @ dms, ao) -> dms_bar = ao.T diag(rho_bar) ao
        D = ao.T @ (ao * rho_bar[:, None])
    if not hermi:
        D = 0.5 * (D + D.T)
    return D, p_bar


def eval_rho_vjp(ao, rho_bar, xctype="LDA", hermi=0):
    """Reverse of ``eval_rho`` w.r.t. dm (density loss, trainer_legacy_no_jit.py:272-275)."""
    if xctype.upper() in ("LDA", "HF"):
        D = ao.T @ (ao * rho_bar[:, None])
    else:
        f = np.array([1.0, 2.0, 2.0, 2.0])
        t = np.einsum("c,cg,cgj->gj", f, rho_bar, ao)
        D = ao[0].T @ t
    if not hermi:
        D = 0.5 * (D + D.T)
    return D
