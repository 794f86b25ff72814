"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's incore
Coulomb / exchange build and its reverse mode.

Follows `_dot_eri_dm_s1` (qedft/train/td/hf_legacy.py:275-286) -- two jnp.einsum calls on the
dense s1 tensor -- and `dot_eri_dm` (:289-299), whose other branch (`eri.size != nao**4`) calls
pyscfad's `_vhf.incore` (third-party, absent) and is not restated.  The reverse mode is what
jax.vjp of those einsums is by definition: the transposed contractions.

Parity unpinned: jax cannot be imported here, and the reference holds no golden vectors for this
function; np.einsum with the reference's own subscript strings is the anchor.
"""
import numpy as np


def dot_eri_dm_s1(eri, dm, with_j=True, with_k=True):
    """hf_legacy.py:275-286 (same subscripts, same reshapes)."""
    dm = np.asarray(dm, dtype=np.float64)
    nao = dm.shape[-1]
    eri = np.asarray(eri, dtype=np.float64).reshape((nao,) * 4)
    dms = dm.reshape(-1, nao, nao)
    vj = vk = None
    if with_j:
        vj = np.einsum("ijkl,xji->xkl", eri, dms).reshape(dm.shape)
    if with_k:
        vk = np.einsum("ijkl,xjk->xil", eri, dms).reshape(dm.shape)
    return vj, vk


def dot_eri_dm(eri, dm, hermi=0, with_j=True, with_k=True):
    """hf_legacy.py:289-299: only the dense-s1 branch exists without pyscfad."""
    dm = np.asarray(dm)
    nao = dm.shape[-1]
    if np.size(eri) != nao**4:
        raise NotImplementedError("packed (s4/s8) ERI goes through pyscfad's _vhf.incore in the reference")
    return dot_eri_dm_s1(eri, dm, with_j, with_k)


def dot_eri_dm_s1_vjp(eri, vj_bar=None, vk_bar=None):
    """Cotangent of dm for cotangents of (vj, vk): transposes of the two einsums above."""
    ref = vj_bar if vj_bar is not None else vk_bar
    ref = np.asarray(ref, dtype=np.float64)
    nao = ref.shape[-1]
    eri = np.asarray(eri, dtype=np.float64).reshape((nao,) * 4)
    out = np.zeros((int(ref.size // (nao * nao)), nao, nao))
    if vj_bar is not None:
        out += np.einsum("ijkl,xkl->xji", eri, np.asarray(vj_bar, dtype=np.float64).reshape(-1, nao, nao))
    if vk_bar is not None:
        out += np.einsum("ijkl,xil->xjk", eri, np.asarray(vk_bar, dtype=np.float64).reshape(-1, nao, nao))
    return out.reshape(ref.shape)


def energy_coulomb(dm, vj):
    """rks_legacy.py:122 `ecoul = einsum("ij,ji", dm, vj) * 0.5`."""
    return 0.5 * np.einsum("ij,ji", dm, vj)


def synthetic_eri(nao, seed=0, symmetric=True):
    """A seeded dense tensor; `symmetric` imposes the 8-fold permutational symmetry of real ERIs
    (built as a Gram form over an auxiliary index, so it is also positive like a physical tensor)."""
    rng = np.random.default_rng(seed)
    if not symmetric:
        return rng.standard_normal((nao,) * 4)
    naux = 2 * nao
    L = rng.standard_normal((naux, nao, nao)) / np.sqrt(naux)
    L = 0.5 * (L + L.transpose(0, 2, 1))
    L = L.reshape(naux, nao * nao)
    return (L.T @ L).reshape((nao,) * 4)
